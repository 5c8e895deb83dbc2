"""GPU: randomised small configurations against the oracle — image counts around the warp / batch boundaries,
one to many classes, every threshold count, empty label and detection files, skewed class priors, ensembles from
empty to everything.  Each case checks TP flags (bit-exact), DCSB (exact) and rewards (explicit ensembles and the
device draw) through the C ABI."""
import numpy as np
import pytest

from helpers import O, flat_tp, oracle_cache
from orie_b200 import data, synth
from orie_b200.synth import DetectorShape

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    M = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 97, 150]))
    C = int(rng.choice([1, 2, 5, 20, 80, 257]))
    T = int(rng.choice([1, 2, 10, 16]))
    weak = DetectorShape(float(rng.uniform(0.2, 0.9)), float(rng.uniform(0.02, 0.12)), float(rng.uniform(1, 40)), int(rng.choice([5, 60, 300])))
    strong = DetectorShape(float(rng.uniform(0.5, 1.0)), float(rng.uniform(0.01, 0.06)), float(rng.uniform(1, 30)), int(rng.choice([5, 60, 300])))
    ds = synth.generate(M, C, float(rng.uniform(0.5, 9)), float(rng.choice([0.0, 0.1, 0.6])), weak, strong, seed,
                        zipf=float(rng.choice([0.0, 1.2])), empty_det_frac=float(rng.choice([0.0, 0.2, 0.7])))
    N = int(rng.choice([0, 1, M // 3, M - 1, M + 5]))
    iouv = np.sort(rng.uniform(0.3, 0.97, size=T)) if T not in (1, 10) else (O.IOU_05 if T == 1 else O.IOU_05_095)
    return ds, M, C, T, N, iouv


@pytest.mark.parametrize("seed", range(24))
def test_random_small_configurations(seed):
    from orie_b200.engine import Engine, clamp_ensemble
    ds, M, C, T, N, iouv = _case(seed)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    eng = Engine(pk, iouv=iouv)
    wd, sd, lc = oracle_cache(pk, iouv)
    wtp, stp, _, _ = eng.tp_flags()
    assert np.array_equal(wtp, flat_tp(wd, len(pk.w_cls), T)) and np.array_equal(stp, flat_tp(sd, len(pk.s_cls), T))
    assert np.array_equal(eng.dcsb(), O.dcsb_all(wd, sd))
    n = clamp_ensemble(M, N)
    em = O.ensemble_matrix(M, N, 77 + seed)
    assert em.shape == (M, n)
    got = eng.orie(N, ens_matrix=em)
    want = O.orie_all(wd, sd, lc, em)
    assert np.abs(got - want).max() < 1e-9, (M, C, T, N)
    # device-drawn ensembles: n distinct images, never the target; rewards follow from the drawn sets
    bits = eng.sample_bits(N, seed=seed)
    member = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(M, -1)[:, :M].astype(bool)
    assert (member.sum(1) == n).all() and not member[np.arange(M), np.arange(M)].any()
    got = eng.orie(N, seed=seed)
    want = O.orie_all(wd, sd, lc, [np.nonzero(member[i])[0] for i in range(M)])
    assert np.abs(got - want).max() < 1e-9, (M, C, T, N)
    eng.close()
