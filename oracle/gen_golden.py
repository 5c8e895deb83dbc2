"""Freeze golden vectors from the LIVE reference (run in the build container).

    python oracle/gen_golden.py

Writes tests/golden/*.npz.  Each fixture holds a small synthetic dataset in
file-content form plus what the unmodified upstream code computed for it:
``box_correct`` TP flags (through ``set_data`` for T=1 and through the
T-generic glue for T=10), ``compute_orie`` with ``np.random.seed(base+idx)``
before every sequential call, and ``compute_dcsb``.  The datasets are written
to a temporary directory in the reference's on-disk formats and read back by
the reference's own loader, so the loader is part of what is pinned.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import orie_b200  # noqa: E402,F401
from orie_b200 import synth  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, config, M, seed, generator kwargs, [(T, N, base_seed)]
    ("coco48", "smoke500", 48, 31, dict(empty_det_frac=0.05), [(1, 12, 500), (10, 12, 500), (10, 0, 7), (1, 1000, 9)]),
    ("voc40", "voc4952", 40, 32, dict(empty_det_frac=0.0, zipf=1.0), [(10, 20, 11), (1, 39, 12)]),
    # the scale-sweep shape: exactly 300 rows per image and detector
    ("dense24", "sweep50k", 24, 33, dict(), [(10, 8, 21), (10, 23, 22), (1, 0, 5)]),
    # duplicate ground-truth boxes (exact IoU ties between same-class labels): upstream's argsort of the candidate
    # pairs (lib/metrics.py:59) is deterministic below 16 elements (insertion sort + reversal -> the HIGHEST label
    # row wins a tie), which is the rule the engine implements; images are kept small so every image stays below 16
    ("ties24", None, 24, 34, dict(), [(10, 6, 41), (1, 23, 42), (10, 0, 43)]),
]


def tied_dataset(M, seed):
    """Small images whose label files repeat rows verbatim (same class, same box), so that a detection has the same
    IoU with two labels.  Fewer than 16 candidate (label, detection) pairs per image and threshold (asserted)."""
    from orie_b200.synth import DetectorShape, Rows
    from oracle import orie_oracle as O
    ds = synth.generate(M, 4, 2.0, 0.0, DetectorShape(.9, .05, 4, 6), DetectorShape(.95, .03, 4, 6), seed)
    rng = np.random.default_rng(seed)
    rows, off = [], [0]
    for i in range(M):
        lab = ds.labels.image(i)
        if len(lab) and i % 4 != 3:                     # three images in four get 1-2 duplicated rows
            k = int(rng.integers(1, 3))
            dup = lab[rng.integers(0, len(lab), size=k)]
            lab = np.concatenate([lab, dup], axis=0)[rng.permutation(len(lab) + k)]
        rows.append(lab)
        off.append(off[-1] + len(lab))
    ds.labels = Rows(np.array(off, dtype=np.int64), np.ascontiguousarray(np.concatenate(rows, axis=0)))
    for det in (ds.weak, ds.strong):                    # the determinism precondition
        for i in range(M):
            d, l = det.image(i), ds.labels.image(i)
            if len(d) and len(l):
                iou = O.pairwise_iou(O.xywh_to_xyxy(l[:, 1:5]), O.xywh_to_xyxy(d[:, 1:5]))
                cand = (iou >= 0.5) & (l[:, :1] == d[None, :, 0])
                assert cand.sum() < 16, (i, int(cand.sum()))
    return ds


def flat(cache, T):
    tp = [c[0].reshape(-1, T) for c in cache]
    return np.concatenate(tp, axis=0) if tp else np.zeros((0, T), dtype=bool)


def main(only=None):
    assert R.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    for name, config, M, seed, kw, runs in CASES:
        if only and name not in only:
            continue
        ds = tied_dataset(M, seed) if config is None else synth.make(config, num_images=M, seed=seed, **kw)
        # make the edge cases explicit: an image without labels but with detections, one with nothing at all
        out = dict(names=np.array(ds.names), num_classes=ds.num_classes,
                   l_off=ds.labels.off, l_rows=ds.labels.rows, w_off=ds.weak.off, w_rows=ds.weak.rows,
                   s_off=ds.strong.off, s_rows=ds.strong.rows)
        with tempfile.TemporaryDirectory() as d:
            w, s, l = synth.write_dirs(ds, d)
            caches = {}
            for T in sorted({r[0] for r in runs}):
                iouv = None if T == 1 else np.linspace(0.5, 0.95, 10)
                caches[T] = R.ref_set_data(w, s, l, iouv)
                wd, sd, lab = caches[T]
                out[f"w_tp_T{T}"] = flat(wd, T)
                out[f"s_tp_T{T}"] = flat(sd, T)
            out["dcsb"] = R.ref_dcsb(*caches[runs[0][0]][:2])
            for T, N, base in runs:
                wd, sd, lab = caches[T]
                out[f"orie_T{T}_N{N}_seed{base}"] = R.ref_orie(wd, sd, lab, N, base)
        out["runs"] = np.array(runs, dtype=np.int64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes")


def main_testmap():
    """Fixture for the realised-mAP sweep (upstream test.py): dataset + CV split + two estimate sets -> test_map."""
    M, K = 48, 3
    ds = synth.make("smoke500", num_images=M, seed=31, empty_det_frac=0.05)
    rng = np.random.default_rng(5)
    fold = rng.permutation(M) % K
    split = np.stack([fold == k for k in range(K)])
    out = dict(l_off=ds.labels.off, l_rows=ds.labels.rows, w_off=ds.weak.off, w_rows=ds.weak.rows,
               s_off=ds.strong.off, s_rows=ds.strong.rows, split=split)
    with tempfile.TemporaryDirectory() as d:
        w, s, l = synth.write_dirs(ds, d)
        est_dirs = []
        for e in range(2):
            ed = os.path.join(d, f"est{e}")
            os.makedirs(ed)
            for k in range(K):
                tr, va = rng.normal(size=int((~split[k]).sum())), rng.normal(size=int(split[k].sum()))
                np.savez(os.path.join(ed, f"estimate{k + 1}.npz"), train_est=tr, val_est=va)
                out[f"est{e}_train{k + 1}"], out[f"est{e}_val{k + 1}"] = tr, va
            est_dirs.append(ed)
        out["test_map"] = R.ref_test_map(w, s, l, est_dirs, split)
    path = os.path.join(OUT, "coco48_testmap.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", out["test_map"].shape)


if __name__ == "__main__":
    if len(sys.argv) > 1:           # python oracle/gen_golden.py dense24  -> only the named fixtures
        main(only=sys.argv[1:])
    else:
        main()
        main_testmap()
