"""GPU: golden vectors frozen from the live reference, the CLI drop-in, the
Python API mirror, edge cases and full-size properties — all through the C ABI."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import O, flat_tp, make_packed, oracle_cache
from orie_b200 import api, data, synth
from orie_b200.synth import Rows

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("dcsb_fit", "rank_norm")) and not p.endswith("_testmap.npz"))


def _engine(pk, iouv, **kw):
    from orie_b200.engine import Engine
    return Engine(pk, iouv=iouv, **kw)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_engine_reproduces_frozen_reference_outputs(path):
    z = np.load(path)
    pk = data.pack(Rows(z["l_off"], z["l_rows"]), Rows(z["w_off"], z["w_rows"]), Rows(z["s_off"], z["s_rows"]))
    M = pk.num_images
    engines = {}
    for T, N, base in z["runs"]:
        T, N, base = int(T), int(N), int(base)
        if T not in engines:
            engines[T] = _engine(pk, O.IOU_05 if T == 1 else O.IOU_05_095)
            wtp, stp, _, _ = engines[T].tp_flags()
            assert np.array_equal(wtp, z[f"w_tp_T{T}"]) and np.array_equal(stp, z[f"s_tp_T{T}"])      # bit-exact
        got = engines[T].orie(N, ens_matrix=api.ensemble_matrix_numpy(M, N, base))
        want = np.where(np.isnan(z[f"orie_T{T}_N{N}_seed{base}"]), 0, z[f"orie_T{T}_N{N}_seed{base}"])
        assert np.abs(got - want).max() <= 1e-6                                                        # north-star tolerance
        assert np.abs(got - want).max() <= 1e-10
    assert np.array_equal(next(iter(engines.values())).dcsb(), z["dcsb"])
    for e in engines.values():
        e.close()


def test_cli_drop_in(tmp_path):
    ds = synth.make("smoke500", num_images=90, seed=21, empty_det_frac=0.05)
    w, s, l = synth.write_dirs(ds, str(tmp_path / "ds"))
    out = str(tmp_path / "out")
    env = dict(os.environ, PYTHONPATH=ROOT)

    def run(*extra):
        subprocess.run([sys.executable, os.path.join(ROOT, "reward.py"), w, s, l, out, *extra], check=True, env=env,
                       capture_output=True, timeout=300)

    pk = data.pack(ds.labels, ds.weak, ds.strong)
    M = pk.num_images
    # orie, default thresholds ([0.5]), numpy ensembles for parity
    run("--method", "orie", "--num-ensemble", "20", "--seed", "5", "--ensembles", "numpy")
    z = np.load(os.path.join(out, "orie20.npz"))
    assert sorted(z.files) == ["reward", "time"] and z["reward"].shape == (M,) and z["reward"].dtype == np.float64
    wd, sd, lc = oracle_cache(pk, O.IOU_05)
    want = O.orie_all(wd, sd, lc, O.ensemble_matrix(M, 20, 5))
    assert np.abs(z["reward"] - want).max() < 1e-9
    # N larger than the dataset: clamped for the maths, file name keeps the typed value (reward.py:91)
    run("--num-ensemble", "5000", "--seed", "6", "--ensembles", "numpy", "--iou-thresholds", "0.5:0.95")
    z = np.load(os.path.join(out, "orie5000.npz"))
    wd10, sd10, lc10 = oracle_cache(pk, O.IOU_05_095)
    want = O.orie_all(wd10, sd10, lc10, O.ensemble_matrix(M, 5000, 6))
    assert np.abs(z["reward"] - want).max() < 1e-9
    # ori == orie with N = 0
    run("--method", "ori")
    z = np.load(os.path.join(out, "orie0.npz"))
    want = O.orie_all(wd, sd, lc, np.zeros((M, 0), dtype=np.int32))
    assert np.abs(z["reward"] - want).max() < 1e-12
    # dcsb
    run("--method", "dcsb")
    z = np.load(os.path.join(out, "dcsb.npz"))
    assert z["reward"].dtype.kind == "i" and np.array_equal(z["reward"], O.dcsb_all(wd, sd))
    # device-drawn ensembles: same seed -> same file content, different seed -> different
    run("--num-ensemble", "30", "--seed", "9"); a = np.load(os.path.join(out, "orie30.npz"))["reward"]
    run("--num-ensemble", "30", "--seed", "9"); b = np.load(os.path.join(out, "orie30.npz"))["reward"]
    run("--num-ensemble", "30", "--seed", "10"); c = np.load(os.path.join(out, "orie30.npz"))["reward"]
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_set_data_mirror_layout(tmp_path):
    ds = synth.make("smoke500", num_images=30, seed=4, empty_det_frac=0.1)
    w, s, l = synth.write_dirs(ds, str(tmp_path))
    wd, sd, labels = api.set_data(w, s, l)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    owd, osd, olc = oracle_cache(pk, O.IOU_05)
    for a, b in zip(wd + sd, owd + osd):
        assert a[0].dtype == bool and a[0].shape == b[0].shape and np.array_equal(a[0], b[0])
        assert np.array_equal(a[1], b[1]) and np.array_equal(pk.class_values[b[2].astype(int)] if len(b[2]) else b[2], a[2])
    for a, b in zip(labels, olc):
        assert len(a) == len(b)


def test_edge_cases():
    from orie_b200._lib import OrieError
    from orie_b200.engine import Engine
    # one image only: every ensemble is empty
    ds = synth.make("smoke500", num_images=1, seed=1)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    eng = Engine(pk, iouv=O.IOU_05_095)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    assert np.abs(eng.orie(1000, seed=1) - O.orie_all(wd, sd, lc, np.zeros((1, 0), dtype=np.int32))).max() < 1e-12
    eng.close()
    # no ground truth anywhere -> NaN upstream -> 0; no detections anywhere -> 0
    ds = synth.make("smoke500", num_images=40, seed=2)
    empty = Rows(np.zeros(41, dtype=np.int64), np.zeros((0, 5)))
    eng = Engine(data.pack(empty, ds.weak, ds.strong), iouv=O.IOU_05)
    assert (eng.orie(10, seed=3) == 0).all()
    eng.close()
    none = Rows(np.zeros(41, dtype=np.int64), np.zeros((0, 6)))
    eng = Engine(data.pack(ds.labels, none, none), iouv=O.IOU_05)
    assert (eng.orie(10, seed=3) == 0).all() and (eng.dcsb() == 0).all()
    eng.close()
    # identical detectors: offloading changes nothing, reward is exactly 0
    eng = Engine(data.pack(ds.labels, ds.weak, ds.weak), iouv=O.IOU_05_095)
    assert (eng.orie(15, seed=4) == 0).all()
    eng.close()
    # 16 thresholds is the limit
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    iou16 = np.linspace(0.2, 0.95, 16)
    eng = Engine(pk, iouv=iou16)
    wd, sd, lc = oracle_cache(pk, iou16)
    em = O.ensemble_matrix(40, 12, 8)
    assert np.abs(eng.orie(12, ens_matrix=em) - O.orie_all(wd, sd, lc, em)).max() < 1e-9
    wtp, _, _, _ = eng.tp_flags()
    assert np.array_equal(wtp, flat_tp(wd, len(pk.w_cls), 16))
    eng.close()
    with pytest.raises(OrieError):
        Engine(pk, iouv=np.linspace(0.1, 0.95, 17))


def test_crowded_image_uses_the_quadratic_matcher():
    # more labels in one image than the shared-memory table holds (ORIE_MAX_LABELS_PER_IMAGE = 4096)
    rng = np.random.default_rng(0)
    G, D = 4200, 300
    lab = np.column_stack([rng.integers(0, 3, G), rng.uniform(0.05, 0.95, (G, 2)), rng.uniform(0.01, 0.05, (G, 2))])
    src = lab[rng.choice(G, D, replace=False)]
    det = np.column_stack([src[:, 0], src[:, 1:3] + rng.normal(0, 0.002, (D, 2)), src[:, 3:5], np.sort(rng.random(D))[::-1]])
    labels = Rows(np.array([0, G, G]), lab)
    dets = Rows(np.array([0, D, D]), det)
    pk = data.pack(labels, dets, dets)
    eng = _engine(pk, O.IOU_05_095)
    wtp, _, wm, _ = eng.tp_flags()
    tp, best, _ = O.match_detections(pk.w_box, pk.w_cls, pk.l_box, pk.l_cls, O.IOU_05_095)
    assert np.array_equal(wtp, tp) and np.array_equal(wm, np.where(tp.any(1), best, -1))
    eng.close()


def test_full_size_coco_shape_properties():
    """COCO-val-shaped 5000 images x 1000-image ensembles, T=10 (BASELINE configs[1])."""
    ds, pk = make_packed("coco5000", M=None)
    M, N = pk.num_images, 1000
    eng = _engine(pk, O.IOU_05_095)
    r1 = eng.orie(N, seed=42)
    # same seed, different wave size and sharding -> identical bits
    r2 = eng.orie(N, seed=42, workspace_budget=eng.workspace_bytes(1024))
    assert np.array_equal(r1, r2)
    a = eng.orie(N, seed=42, t0=1024, nt=2048)
    assert np.array_equal(r1[1024:3072], a)
    # spot parity against the oracle on the device-drawn ensembles
    wtp, stp, _, _ = eng.tp_flags()
    bits = eng.sample_bits(N, seed=42)
    targets = [0, 1, 2499, 4998, 4999]
    member = ((bits[targets][:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(len(targets), -1)[:, :M].astype(bool)
    assert (member.sum(1) == N).all()
    sys.path.insert(0, ROOT)
    import bench
    wd, sd, lc = bench.cpu_cache_from_flags(pk, O.IOU_05_095, wtp, stp)
    for r, i in enumerate(targets):
        want = O.orie_one(i, wd, sd, lc, np.nonzero(member[r])[0])[0]
        assert abs(r1[i] - (0 if np.isnan(want) else want)) < 1e-6
    # TP matching: spot check of images against the CPU matcher
    for i in range(0, M, 499):
        a, b, la, lb = pk.w_off[i], pk.w_off[i + 1], pk.l_off[i], pk.l_off[i + 1]
        tp, _, _ = O.match_detections(pk.w_box[a:b], pk.w_cls[a:b], pk.l_box[la:lb], pk.l_cls[la:lb], O.IOU_05_095)
        assert np.array_equal(wtp[a:b], tp)
    eng.close()
    # identical detectors at full size: exactly zero
    eng = _engine(data.pack(ds.labels, ds.weak, ds.weak), O.IOU_05_095)
    assert (eng.orie(N, seed=1) == 0).all()
    eng.close()


def test_realised_map_sweep_matches_frozen_reference(tmp_path):
    """The reference's test.py (realised mAP vs offloading ratio) through the engine and through the CLI drop-in."""
    from orie_b200 import evaluate
    from test_oracle_golden import _testmap_case
    z, lab, wk, st, masks = _testmap_case(tmp_path)
    got = evaluate.realised_map(lab, wk, st, masks, iouv=O.IOU_05)
    assert np.abs(got.reshape(2, 11) - z["test_map"]).max() < 1e-9
    # every image weak / every image strong == plain dataset mAP of each detector
    pk = data.pack(lab, wk, st)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    gt = np.concatenate(lc).astype(int)
    both = evaluate.realised_map(lab, wk, st, np.stack([np.zeros(48, bool), np.ones(48, bool)]), iouv=O.IOU_05_095)
    for cache, val in ((wd, both[0]), (sd, both[1])):
        cols = [np.concatenate(c, axis=0) for c in zip(*cache)]
        assert abs(np.mean(O.ap_by_class(*cols, gt)) - val) < 1e-9
    # CLI
    ds_dir = tmp_path / "ds"
    from orie_b200.synth import SynthDataset
    names = [f"{i:012d}" for i in range(lab.num_images)]
    w, s, l = synth.write_dirs(SynthDataset(names, lab, wk, st, 80), str(ds_dir))
    np.save(tmp_path / "split.npy", z["split"])
    out = tmp_path / "out"
    subprocess.run([sys.executable, os.path.join(ROOT, "test.py"), w, s, l, str(tmp_path / "split.npy"), str(out),
                    "--estimates", str(tmp_path / "est0"), str(tmp_path / "est1")], check=True, capture_output=True,
                   timeout=300, env=dict(os.environ, PYTHONPATH=ROOT))
    assert np.abs(np.load(out / "test_map.npy") - z["test_map"]).max() < 1e-9
