"""CPU: host-side logic — loader semantics, packer, CLI surface, sharding, C ABI exports."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import orie_b200  # noqa: F401
from orie_b200 import api, data, engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_loader_file_semantics(tmp_path):
    lab, wk = tmp_path / "labels", tmp_path / "weak"
    lab.mkdir(); wk.mkdir()
    (lab / "b.txt").write_text("3 0.5 0.5 0.2 0.2\n7 0.1 0.1 0.1 0.1\n")
    (lab / "a.jpg.txt").write_text("")                       # image without labels; name keeps its inner dot
    (lab / "c.txt").write_text("1 0.3 0.3 0.1 0.1\n")
    names = data.list_images(str(lab))
    assert names == ["a.jpg", "b", "c"]                      # sorted listing, last extension stripped
    (wk / "b.txt").write_text("3 0.5 0.5 0.2 0.2 0.9\n3 0.4 0.4 0.2 0.2 0.25\n")
    np.save(wk / "b.npy", np.zeros((5, 6)))                  # .txt wins over .npy
    np.save(wk / "c.npy", np.array([[1, 0.3, 0.3, 0.1, 0.1, 0.7]]))
    rows = data.read_rows(str(wk), names, True)              # "a.jpg": no file -> no rows
    assert rows.off.tolist() == [0, 0, 2, 3]
    assert rows.rows[:, 5].tolist() == [0.9, 0.25, 0.7]
    labels = data.read_rows(str(lab), names, False)
    assert labels.off.tolist() == [0, 0, 2, 3]
    pk = data.pack(labels, rows, rows)
    assert pk.class_values.tolist() == [1, 3, 7] and pk.l_cls.tolist() == [1, 2, 0]
    assert np.allclose(pk.l_box[0], [0.4, 0.4, 0.6, 0.6])
    (wk / "c.txt").write_text("1 0.3 oops 0.1 0.1 0.7\n")
    with pytest.raises(ValueError):
        data.read_rows(str(wk), names, True)
    (wk / "c.txt").write_text("1 0.3 0.3\n")
    with pytest.raises(ValueError):
        data.read_rows(str(wk), names, True)


def test_write_dirs_round_trip(tmp_path):
    ds = synth.make("voc4952", num_images=25, seed=3, empty_det_frac=0.1)
    w, s, l = synth.write_dirs(ds, str(tmp_path))
    names, lab, wk, st = data.load_dirs(w, s, l)
    assert names == ds.names
    for a, b in ((lab, ds.labels), (wk, ds.weak), (st, ds.strong)):
        assert np.array_equal(a.off, b.off) and np.array_equal(a.rows, b.rows)    # repr() floats round-trip exactly


def test_synthetic_data_is_tie_free():
    ds = synth.make("smoke500", num_images=300)
    for r in (ds.weak, ds.strong):
        assert len(np.unique(r.rows[:, 5])) == len(r.rows)
        for i in range(0, 300, 17):                                                # file order = confidence descending
            c = r.image(i)[:, 5]
            assert (np.diff(c) < 0).all()
    assert len(np.unique(ds.labels.rows[:, 1:], axis=0)) == len(ds.labels.rows)


def test_cli_surface_and_output_layout(tmp_path):
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("orie_cli", os.path.join(ROOT, "reward.py"))
    cli = importlib.util.module_from_spec(spec); spec.loader.exec_module(cli)
    o = cli.getargs(["w", "s", "l", "out"])
    assert (o.method, o.num_ensemble, o.iou_thresholds, o.ensembles) == ("orie", 1000, "0.5", "device")
    assert cli.getargs(["w", "s", "l", "out", "--method", "ori"]).method == "ori"
    with pytest.raises(SystemExit):
        cli.getargs(["w", "s", "l", "out", "--method", "nope"])
    p = api.save_rewards(str(tmp_path / "deep" / "dir"), "orie", 5000, np.arange(3.0), 1.5)
    z = np.load(p)
    assert os.path.basename(p) == "orie5000.npz" and sorted(z.files) == ["reward", "time"]
    assert z["reward"].dtype == np.float64 and z["time"].shape == () and float(z["time"]) == 1.5
    assert os.path.basename(api.save_rewards(str(tmp_path), "ori", 77, np.zeros(2), 0.0)) == "orie0.npz"
    p = api.save_rewards(str(tmp_path), "dcsb", 1000, np.array([1, -2]), 0.1)
    assert os.path.basename(p) == "dcsb.npz" and np.load(p)["reward"].dtype.kind == "i"
    assert api.parse_iou_thresholds("0.5").tolist() == [0.5]
    assert np.array_equal(api.parse_iou_thresholds("0.5:0.95"), np.linspace(0.5, 0.95, 10))


def test_numpy_ensembles_follow_the_reference_recipe():
    from oracle import orie_oracle as O
    M, N = 37, 9
    em = api.ensemble_matrix_numpy(M, N, 123)
    assert np.array_equal(em, O.ensemble_matrix(M, N, 123))
    np.random.seed(123 + 4)                                   # what upstream would draw after np.random.seed(base+idx)
    idx = np.arange(M - 1); idx[4:] += 1
    assert np.array_equal(em[4], np.random.permutation(idx)[:N])
    assert api.ensemble_matrix_numpy(M, 1000, 1).shape == (M, M - 1)
    assert engine.clamp_ensemble(10, -3) == 0 and engine.clamp_ensemble(10, 50) == 9


def test_shard_ranges_cover_the_targets_once():
    for M in (1, 31, 32, 33, 500, 4952, 5000, 50000):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(M, dtype=int)
            for r in range(world):
                t0, nt = engine.shard_range(M, r, world)
                assert (nt == 0 or t0 % 32 == 0) and nt >= 0 and t0 + nt <= M
                seen[t0:t0 + nt] += 1
            assert (seen == 1).all()
            per = engine.shard_range(M, 0, world)[1]
            assert all(engine.shard_range(M, r, world)[1] <= per for r in range(world))


def test_shard_mode_follows_the_dataset_size():
    assert engine.pick_shard(5000) == "classes" and engine.pick_shard(50000) == "classes"
    assert engine.shard_plan(5000, 8) == (8, 1) and engine.shard_plan(50000, 8) == (8, 1)
    assert engine.shard_plan(5000, 8, num_classes=3) == (2, 4) and engine.shard_plan(5000, 8, num_classes=1) == (1, 8)
    assert engine.pick_shard(50000, "classes") == "classes" and engine.pick_shard(10, "targets") == "targets"


def test_c_abi_exports_every_declared_symbol():
    from orie_b200 import _lib
    header = open(os.path.join(ROOT, "include", "orie_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)             # prototypes only, not prose
    declared = set(re.findall(r"\b(orie_[a-z_]+)\s*\(", header))
    assert declared, "no prototypes found in include/orie_b200.h"
    lib = ctypes.CDLL(_lib.lib_path()) if os.path.exists(_lib.lib_path()) else _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"liborie_b200.so does not export {name}"
    assert set(_lib.SYMBOLS) <= declared
    lib.orie_version.restype = ctypes.c_int
    assert lib.orie_version() >= 100
    # argument validation runs on the host without touching the GPU
    lib.orie_last_error.restype = ctypes.c_char_p
    rc = lib.orie_match(None, None, None, None, None, None, None, 99, ctypes.c_int64(1), None, None, None, None)
    assert rc == 3 and b"T=99" in lib.orie_last_error()


def test_ctypes_structs_match_the_header():
    """orie_tuning_t and orie_index_info_t cross the C ABI by value layout: the ctypes mirrors in _lib.py must list the
    header's fields in the header's order with the header's types."""
    import ctypes as C
    from orie_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "orie_b200.h")).read()
    ctype = {"int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double}

    def fields(name):
        body = re.search(r"typedef struct \{([^{}]*)\} " + name + ";", hdr).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, names = decl.split(None, 1)
            out += [(n.strip(), ctype[ty]) for n in names.split(",")]
        return out

    assert fields("orie_tuning_t") == list(_lib.Tuning._fields_)
    info_cls = next(v for v in vars(_lib).values() if isinstance(v, type) and issubclass(v, C.Structure)
                    and any(f[0] == "num_images" for f in getattr(v, "_fields_", [])))
    assert fields("orie_index_info_t") == list(info_cls._fields_)


def test_integration_stub_matches_the_header():
    """INTEGRATION.md shows the ctypes stub a maintainer of the reference would add; every ``argtypes`` list in it must
    have as many entries as the header's prototype has parameters."""
    import ctypes as C
    hdr = open(os.path.join(ROOT, "include", "orie_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {m.group(1): m.group(2) for m in re.finditer(r"\b(orie_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", hdr, re.S)}
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    env = {"C": C, "_vp": C.c_void_p, "_i64": C.c_int64}
    seen = 0
    for m in re.finditer(r"_lib\.(orie_[a-z_0-9]+)\.argtypes = (\[.*?\])(?:;|\s*#|\s*$)", doc, re.M):
        name, expr = m.group(1), m.group(2)
        assert name in protos, name
        params = [a for a in protos[name].split(",") if a.strip() and a.strip() != "void"]
        assert len(eval(expr, env)) == len(params), (name, expr, protos[name])
        seen += 1
    assert seen >= 8


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ds = synth.make("smoke500", num_images=8)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.Engine(pk)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "edgeml-object-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
    text = open(os.path.join(ROOT, "reward.py")).read()
    assert "oracle" not in text


def test_class_partition_and_reward_formula():
    ds = synth.make("smoke500", num_images=80, seed=6, zipf=1.0)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    for world in (1, 2, 3, 8, 200):
        owner = engine.class_partition(pk, world)
        assert owner.min() >= 0 and owner.max() < world and len(owner) == pk.num_classes
        rows = 0
        for r in range(world):
            sh = engine.class_shard(pk, r, world)
            assert sh.num_images == pk.num_images and len(sh.w_off) == pk.num_images + 1
            assert set(sh.class_values.tolist()) == set(pk.class_values[owner == r].tolist())
            rows += len(sh.w_cls) + len(sh.s_cls) + len(sh.l_cls)
            if len(sh.w_cls):                                     # rows keep their file order inside an image
                i = int(np.argmax(np.diff(sh.w_off)))
                keep = owner[pk.w_cls[pk.w_off[i]:pk.w_off[i + 1]]] == r
                assert np.array_equal(sh.w_conf[sh.w_off[i]:sh.w_off[i + 1]], pk.w_conf[pk.w_off[i]:pk.w_off[i + 1]][keep])
        assert rows == len(pk.w_cls) + len(pk.s_cls) + len(pk.l_cls)
    sums = np.array([[1.0, 2.0, 4.0], [3.0, 1.0, 2.0], [0.0, 0.0, 0.0]])
    r = engine.rewards_from_sums(sums, 10, 5)
    assert np.allclose(r, [(2 / 40 - 1 / 40) * 6, (1 / 20 - 3 / 20) * 6, 0.0])
