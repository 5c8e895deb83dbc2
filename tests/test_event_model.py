"""CPU: the engine's formulation (global sort + membership + reverse AP sweep,
oracle/event_model.py) against the straightforward oracle."""
import numpy as np
import pytest

from helpers import O, flat_tp, make_packed, oracle_cache
from oracle import event_model as E


def masks(cache, D, T):
    tp = flat_tp(cache, D, T).astype(np.int64)
    return (tp << np.arange(T)).sum(axis=1)


@pytest.mark.parametrize("M,N,T,segch,seed", [(36, 10, 1, 1, 1), (36, 12, 10, 2, 2), (24, 0, 10, 4, 3), (30, 29, 10, 1, 4)])
def test_event_model_equals_oracle(M, N, T, segch, seed):
    iouv = O.IOU_05 if T == 1 else O.IOU_05_095
    _, pk = make_packed(M=M, seed=seed, empty_det_frac=0.06)
    wd, sd, lc = oracle_cache(pk, iouv)
    ix = E.build_index(M, pk.num_classes, pk.w_off, pk.w_cls.astype(np.int64), pk.w_conf, masks(wd, len(pk.w_cls), T),
                       pk.s_off, pk.s_cls.astype(np.int64), pk.s_conf, masks(sd, len(pk.s_cls), T),
                       pk.l_off, pk.l_cls.astype(np.int64), seg_chunks=segch)
    em = O.ensemble_matrix(M, N, 7)
    want = O.orie_all(wd, sd, lc, em)
    got = np.array([E.reward_target(ix, j, em[j], T) for j in range(M)])
    assert np.abs(got - want).max() < 1e-9
    # difference-only sweep with early exit (what ap_kernel<false> does): same rewards, far fewer steps
    delta = [E.reward_target_delta(ix, j, em[j], T) for j in range(M)]
    assert np.abs(np.array([d[0] for d in delta]) - want).max() < 1e-9


@pytest.mark.parametrize("seed", range(6))
def test_difference_sweep_exits_on_random_shapes(seed):
    """The difference-only sweep (early exits of ap_kernel<false>, including the shared-state claim its tail loop relies
    on, asserted inside the model) on random small shapes: few classes so that sweeps are deep, skewed priors, empty
    files, every ensemble size."""
    rng = np.random.default_rng(50 + seed)
    M = int(rng.choice([5, 20, 33, 48]))
    N = int(rng.choice([0, 1, M // 2, M - 1]))
    from orie_b200 import data, synth
    from orie_b200.synth import DetectorShape
    ds = synth.generate(M, int(rng.choice([1, 2, 6])), float(rng.uniform(1, 6)), float(rng.choice([0.0, 0.3])),
                        DetectorShape(float(rng.uniform(.3, .9)), .08, float(rng.uniform(2, 25)), 60),
                        DetectorShape(float(rng.uniform(.5, 1.)), .04, float(rng.uniform(2, 25)), 60), 500 + seed,
                        zipf=float(rng.choice([0.0, 1.0])), empty_det_frac=float(rng.choice([0.0, 0.3])))
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    T, iouv = 10, O.IOU_05_095
    wd, sd, lc = oracle_cache(pk, iouv)
    ix = E.build_index(M, pk.num_classes, pk.w_off, pk.w_cls.astype(np.int64), pk.w_conf, masks(wd, len(pk.w_cls), T),
                       pk.s_off, pk.s_cls.astype(np.int64), pk.s_conf, masks(sd, len(pk.s_cls), T),
                       pk.l_off, pk.l_cls.astype(np.int64), seg_chunks=int(rng.choice([1, 3])))
    em = O.ensemble_matrix(M, N, 3)
    want = O.orie_all(wd, sd, lc, em)
    got = np.array([E.reward_target_delta(ix, j, em[j], T)[0] for j in range(M)])
    assert np.abs(got - want).max() < 1e-9


def test_integer_recall_rule_equals_float_comparison():
    # "x_g >= fl(k / n_l)" decided in integers (oracle/event_model.py:grid_lo) for every small case
    for n_l in list(range(1, 260)) + [1000, 4096, 99991]:
        ks = range(0, n_l + 1) if n_l < 300 else list(range(0, 200)) + [n_l // 2, n_l - 1, n_l]
        for k in ks:
            want = int(np.searchsorted(E.GRID, k / n_l, side="left"))   # first g with x_g >= k/n_l
            assert E.grid_lo(k * 100, n_l) == want, (k, n_l)
