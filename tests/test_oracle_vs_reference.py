"""CPU, build container only: the oracle and the loaders against the LIVE
reference (skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from helpers import O, make_packed, records
from oracle import ref_harness as R
from orie_b200 import data, synth

pytestmark = pytest.mark.skipif(not R.available(), reason="live reference not mounted")


@pytest.fixture(scope="module")
def dirs(tmp_path_factory):
    ds = synth.make("smoke500", num_images=40, seed=5, empty_det_frac=0.08)
    root = tmp_path_factory.mktemp("ds")
    return ds, synth.write_dirs(ds, str(root))


def test_loader_matches_reference_loader(dirs):
    ds, (w, s, l) = dirs
    _, _, rdata = R.modules()
    names = data.list_images(l)
    assert names == ds.names
    mine = data.pack(*[data.read_rows(p, names, c) for p, c in ((l, False), (w, True), (s, True))])
    ref_w, ref_s, ref_l = rdata.load_data(w, names, True), rdata.load_data(s, names, True), rdata.load_data(l, names)
    W, S, L = records(mine)
    for a, b in zip(W + S, ref_w + ref_s):
        assert len(a) == len(b)
        if len(a):
            assert np.array_equal(mine.class_values[a[0]], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for a, b in zip(L, ref_l):
        assert len(a) == len(b)
        if len(a):
            assert np.array_equal(mine.class_values[a[0]], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("T", [1, 10])
def test_cache_and_rewards_match_reference(dirs, T):
    ds, (w, s, l) = dirs
    iouv = None if T == 1 else O.IOU_05_095
    rw, rs, rl = R.ref_set_data(w, s, l, iouv)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    W, S, L = records(pk)
    wd, sd, lc = O.build_cache(W, S, L, O.IOU_05 if T == 1 else O.IOU_05_095)
    for a, b in zip(wd + sd, rw + rs):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for N, base in ((8, 100), (0, 5), (1000, 9)):
        want = R.ref_orie(rw, rs, rl, N, base)
        em = O.ensemble_matrix(len(lc), N, base)
        got = np.array([O.orie_one(i, wd, sd, lc, em[i])[0] for i in range(len(lc))])
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want)) <= 1e-12
    assert np.array_equal(R.ref_dcsb(rw, rs), O.dcsb_all(wd, sd))


def test_box_correct_random_cases():
    rng = np.random.default_rng(1)
    iouv = O.IOU_05_095
    for _ in range(200):
        n, m = int(rng.integers(1, 20)), int(rng.integers(1, 8))
        det = rng.uniform(0, 1, (n, 2)); det = np.concatenate([det, det + rng.uniform(0.05, 0.4, (n, 2))], axis=1)
        lab = det[rng.integers(0, n, m)] + rng.normal(0, 0.02, (m, 4))
        dc, lc = rng.integers(0, 3, n), rng.integers(0, 3, m)
        conf = rng.random(n)
        want = R.ref_box_correct(np.column_stack([det, conf, dc]), np.column_stack([lc, lab]), iouv)
        got, _, _ = O.match_detections(det, dc, lab, lc, iouv)
        assert np.array_equal(got, want)


def test_box_correct_with_duplicate_labels_iou_ties():
    """Exact IoU ties (a label row repeated verbatim): upstream's argsort (lib/metrics.py:59) is deterministic below 16
    candidate pairs, and the oracle's rule (ties -> highest label row, the engine's match.cu rule) reproduces its flags."""
    rng = np.random.default_rng(7)
    iouv = O.IOU_05_095
    checked = 0
    for _ in range(400):
        n, m = int(rng.integers(1, 5)), int(rng.integers(1, 3))
        det = rng.uniform(0, 1, (n, 2)); det = np.concatenate([det, det + rng.uniform(0.05, 0.4, (n, 2))], axis=1)
        lab = det[rng.integers(0, n, m)] + rng.normal(0, 0.02, (m, 4))
        lc = rng.integers(0, 2, m)
        k = int(rng.integers(1, 3))                       # repeat 1-2 label rows verbatim, shuffled into the file
        pick = rng.integers(0, m, k)
        perm = rng.permutation(m + k)
        lab, lc = np.concatenate([lab, lab[pick]])[perm], np.concatenate([lc, lc[pick]])[perm]
        dc = rng.integers(0, 2, n)
        conf = rng.random(n)
        if ((O.pairwise_iou(lab, det) >= 0.5) & (lc[:, None] == dc[None, :])).sum() >= 16:
            continue
        want = R.ref_box_correct(np.column_stack([det, conf, dc]), np.column_stack([lc, lab]), iouv)
        got, _, _ = O.match_detections(det, dc, lab, lc, iouv)
        assert np.array_equal(got, want)
        checked += 1
    assert checked > 300
