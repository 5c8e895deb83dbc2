"""Shared test helpers: packed dataset -> oracle records, TP masks, etc."""
import numpy as np

import orie_b200  # noqa: F401  (import shim)
from orie_b200 import data, synth
from oracle import orie_oracle as O


def records(pk):
    """Packed arrays -> the reference loader's list-of-tuples (already xyxy)."""
    M = pk.num_images

    def recs(off, cls, box, conf):
        out = []
        for i in range(M):
            a, b = off[i], off[i + 1]
            if a == b:
                out.append(())
            elif conf is not None:
                out.append((cls[a:b].astype(np.int64), box[a:b], conf[a:b]))
            else:
                out.append((cls[a:b].astype(np.int64), box[a:b]))
        return out

    return (recs(pk.w_off, pk.w_cls, pk.w_box, pk.w_conf), recs(pk.s_off, pk.s_cls, pk.s_box, pk.s_conf),
            recs(pk.l_off, pk.l_cls, pk.l_box, None))


def oracle_cache(pk, iouv):
    W, S, L = records(pk)
    return O.build_cache(W, S, L, iouv)


def flat_tp(cache, D, T):
    out = np.zeros((D, T), dtype=bool)
    pos = 0
    for tp, _, _ in cache:
        out[pos:pos + len(tp)] = tp
        pos += len(tp)
    assert pos == D
    return out


def make_packed(config="smoke500", M=200, seed=None, **kw):
    """M=None keeps the configuration's own image count."""
    ds = synth.make(config, num_images=M, seed=seed, **kw)
    return ds, data.pack(ds.labels, ds.weak, ds.strong)
