"""CPU: the oracle restatement against the golden vectors frozen from the live
reference (tests/golden/*.npz, generator oracle/gen_golden.py)."""
import glob
import os

import numpy as np
import pytest

from helpers import O, flat_tp, oracle_cache
from orie_b200 import data
from orie_b200.synth import Rows

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("dcsb_fit", "rank_norm")) and not p.endswith("_testmap.npz"))


def load_case(path):
    z = np.load(path, allow_pickle=False)
    lab, wk, st = Rows(z["l_off"], z["l_rows"]), Rows(z["w_off"], z["w_rows"]), Rows(z["s_off"], z["s_rows"])
    return z, data.pack(lab, wk, st)


def iouv_for(T):
    return O.IOU_05 if T == 1 else O.IOU_05_095


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_frozen_reference_outputs(path):
    z, pk = load_case(path)
    M = pk.num_images
    assert len(GOLDEN) >= 2
    caches = {}
    for T, N, base in z["runs"]:
        T, N, base = int(T), int(N), int(base)
        if T not in caches:
            caches[T] = oracle_cache(pk, iouv_for(T))
            wd, sd, _ = caches[T]
            assert np.array_equal(flat_tp(wd, len(pk.w_cls), T), z[f"w_tp_T{T}"])
            assert np.array_equal(flat_tp(sd, len(pk.s_cls), T), z[f"s_tp_T{T}"])
        wd, sd, lc = caches[T]
        em = O.ensemble_matrix(M, N, base)
        got = np.array([O.orie_one(i, wd, sd, lc, em[i])[0] for i in range(M)])
        want = z[f"orie_T{T}_N{N}_seed{base}"]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want)) <= 1e-12
    wd, sd, _ = next(iter(caches.values()))
    assert np.array_equal(O.dcsb_all(wd, sd), z["dcsb"])


def test_two_matching_restatements_agree():
    rng = np.random.default_rng(0)
    for trial in range(300):
        n, m = rng.integers(0, 12), rng.integers(0, 6)
        det = rng.uniform(0, 1, (n, 2)); det = np.concatenate([det, det + rng.uniform(0.05, 0.5, (n, 2))], axis=1)
        lab = rng.uniform(0, 1, (m, 2)); lab = np.concatenate([lab, lab + rng.uniform(0.05, 0.5, (m, 2))], axis=1)
        dc, lc = rng.integers(0, 2, n), rng.integers(0, 2, m)
        tp, best, _ = O.match_detections(det, dc, lab, lc, O.IOU_05_095 - 0.45)
        tp2, pairs = O.match_detections_sortunique(det, dc, lab, lc, O.IOU_05_095 - 0.45)
        assert np.array_equal(tp, tp2)
        for t, pr in enumerate(pairs):
            assert {(int(best[d]), int(d)) for d in np.nonzero(tp[:, t])[0]} == {(int(a), int(b)) for a, b in pr}


def test_known_answers_from_survey():
    # no fallback to the second-best label
    labs = np.array([[0, 0, 10, 10], [0, 1, 11, 10]], dtype=float)
    dets = np.array([[0, 0, 10, 10], [0.2, 0, 10.2, 10]], dtype=float)
    tp, _, _ = O.match_detections(dets, np.zeros(2, int), labs, np.zeros(2, int), O.IOU_05)
    assert tp[:, 0].tolist() == [True, False]
    # np.interp takes the LAST knot on an exact hit
    assert np.interp([0.5], [0, .5, .5, .5, 1], [1, .9, .8, .7, 0])[0] == 0.7
    # an image without labels and an empty ensemble -> NaN upstream -> 0
    wd = [(np.zeros((1, 1), bool), np.array([0.9]), np.array([0]))]
    out = O.orie_all(wd, wd, [np.array([])], np.zeros((1, 0), dtype=np.int32))
    assert out[0] == 0


def _testmap_case(tmp_path):
    from orie_b200 import evaluate
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "coco48_testmap.npz"))
    lab, wk, st = Rows(z["l_off"], z["l_rows"]), Rows(z["w_off"], z["w_rows"]), Rows(z["s_off"], z["s_rows"])
    split = z["split"]
    dirs = []
    for e in range(2):
        d = tmp_path / f"est{e}"
        d.mkdir()
        for k in range(split.shape[0]):
            np.savez(d / f"estimate{k + 1}.npz", train_est=z[f"est{e}_train{k + 1}"], val_est=z[f"est{e}_val{k + 1}"])
        dirs.append(str(d))
    masks = np.concatenate([evaluate.offload_masks_from_estimates(d, split) for d in dirs], axis=0)
    return z, lab, wk, st, masks


def test_offload_policy_and_oracle_reproduce_frozen_test_map(tmp_path):
    """test.py:26-42 — the threshold policy (host code of the product) and the oracle's dataset-wide AP against the
    test_map the live reference produced."""
    z, lab, wk, st, masks = _testmap_case(tmp_path)
    assert masks.shape == (22, lab.num_images)
    assert not masks[0].any() or masks[0].sum() < masks[10].sum()
    pk = data.pack(lab, wk, st)
    wd, sd, lc = oracle_cache(pk, O.IOU_05)
    gt = np.concatenate(lc).astype(int)
    got = []
    for m in masks:
        recs = [sd[i] if m[i] else wd[i] for i in range(len(m))]
        cols = [np.concatenate(c, axis=0) for c in zip(*recs)]
        got.append(np.mean(O.ap_by_class(*cols, gt)))
    assert np.abs(np.array(got).reshape(2, 11) - z["test_map"]).max() < 1e-12
