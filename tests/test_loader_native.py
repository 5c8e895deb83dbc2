"""CPU: the native file reader (csrc/loader.cpp through include/orie_io.h) against the Python statement of
lib/data.py:11-43 and, in the build container, against the live reference loader itself."""
import ctypes
import os
import re

import numpy as np
import pytest

import orie_b200  # noqa: F401
from orie_b200 import _io, data, synth
from oracle import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same_rows(a, b):
    return np.array_equal(a.off, b.off) and a.rows.shape == b.rows.shape and \
        np.array_equal(a.rows.view(np.uint64), b.rows.view(np.uint64))        # bit for bit, NaN-safe


def test_io_abi_exports_every_declared_symbol():
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "orie_io.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(orie_io_[a-z_]+)\s*\(", header))
    lib = _io.load()
    assert declared and declared == set(_io.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    rc = lib.orie_io_read_rows(None, None, 0, 0, 0, ctypes.byref(ctypes.c_void_p(0)))
    assert rc == 1 and b"bad arguments" in lib.orie_io_last_error()


@pytest.mark.parametrize("config,M", [("smoke500", 300), ("voc4952", 120)])
def test_native_reader_is_bit_identical_to_the_python_reader(tmp_path, config, M):
    ds = synth.make(config, num_images=M, seed=11, empty_det_frac=0.05)
    w, s, l = synth.write_dirs(ds, str(tmp_path))             # weak / labels as repr() text, strong as .npy
    names = data.list_images(l)
    for path, conf, want in ((l, False, ds.labels), (w, True, ds.weak), (s, True, ds.strong)):
        py = data.read_rows(path, names, conf, native=False)
        for threads in (1, 3, 0):
            assert same_rows(data.read_rows(path, names, conf, native=True, threads=threads), py)
        assert same_rows(py, want)


def test_native_reader_number_formats_and_layout(tmp_path):
    d = tmp_path / "dets"
    d.mkdir()
    names = ["a", "b", "c", "d", "e", "f", "g"]
    (d / "a.txt").write_text("3 0.5 5e-1 .25 1. 0.123456789012345678\r\n"             # CRLF, exponent, bare dots
                             "  7 1e-320 0.1 0.2 0.3 9.999999999999999e-01  \n"        # denormal, padded line
                             "0 0.1 0.2 0.3 0.4 0.5 0.6 0.75")                          # extra columns, no newline
    (d / "b.txt").write_text("")                                                        # empty .txt beats b.npy
    np.save(d / "b.npy", np.ones((4, 6)))
    np.save(d / "c.npy", np.array([[1, .1, .2, .3, .4, .9], [2, .5, .5, .1, .1, .8]], dtype=np.float32))
    np.save(d / "d.npy", np.zeros((0, 6)))
    np.save(d / "e.npy", np.array([[5, .1, .2, .3, .4, .6, .7]]))                       # 7 columns: last is the confidence
    np.save(d / "f.npy", np.array([[1, 0, 0, 1, 1, 1]], dtype=np.int64))                # not f4/f8: Python re-reads it
    # "g": no file at all
    py = data.read_rows(str(d), names, True, native=False)
    nat = data.read_rows(str(d), names, True, native=True)
    assert same_rows(nat, py)
    assert py.off.tolist() == [0, 3, 3, 5, 5, 6, 7, 7]
    # zip(*rows) semantics: table cut to the shortest row (6 tokens), so row 3's "confidence" is its 6th token
    assert py.rows[:3, 5].tolist() == [0.123456789012345678, 9.999999999999999e-01, 0.5]
    assert py.rows[1, 1] == 1e-320 and py.rows[5, 5] == 0.7 and py.rows[6].tolist() == [1, 0, 0, 1, 1, 1]
    off, rows, fb = _io.read_rows(str(d), names, True)
    assert fb.tolist() == [5] and off.tolist() == [0, 3, 3, 5, 5, 6, 6, 6]             # only the int64 .npy was handed back


@pytest.mark.parametrize("text", ["1 0.3 oops 0.1 0.1 0.7\n", "1 0.3  0.3 0.1 0.1 0.7\n", "1 0.3 0.3 0.1 0.1 0.7\n\n",
                                  "1 0.3 0.3\n", "1\t0.3 0.3 0.1 0.1 0.7\n", "0x10 0.3 0.3 0.1 0.1 0.7\n"])
def test_files_upstream_rejects_still_raise(tmp_path, text):
    d = tmp_path / "dets"
    d.mkdir()
    (d / "x.txt").write_text(text)
    (d / "y.txt").write_text("1 0.3 0.3 0.1 0.1 0.7\n")
    with pytest.raises(ValueError):
        data.read_rows(str(d), ["x", "y"], True, native=False)
    with pytest.raises(ValueError):
        data.read_rows(str(d), ["x", "y"], True, native=True)
    off, rows, fb = _io.read_rows(str(d), ["x", "y"], True)
    assert fb.tolist() == [0] and off.tolist() == [0, 0, 1]


def test_spellings_only_python_accepts_go_through_the_fallback(tmp_path):
    d = tmp_path / "dets"
    d.mkdir()
    (d / "x.txt").write_text("1 0.3 0.3 0.1 0.1 nan\n2 0.3 0.3 0.1 0.1 inf\n3 0.3 0.3 0.1 0.1 1_0\n")
    (d / "y.txt").write_text("1 0.3 0.3 0.1 0.1 0.7\n")
    py = data.read_rows(str(d), ["y", "x"], True, native=False)
    nat = data.read_rows(str(d), ["y", "x"], True, native=True)
    assert same_rows(nat, py) and py.off.tolist() == [0, 1, 4]


@pytest.mark.skipif(not ref_harness.available(), reason="live reference not mounted (GPU box)")
def test_native_reader_matches_the_live_reference_loader(tmp_path):
    ds = synth.make("smoke500", num_images=60, seed=5, empty_det_frac=0.1)
    w, s, l = synth.write_dirs(ds, str(tmp_path))
    names = data.list_images(l)
    _, _, ref_data = ref_harness.modules()
    for path, conf in ((l, False), (w, True), (s, True)):
        ref = ref_data.load_data(path, names, conf)            # list of tuples: (cls int, xyxy, [conf]) or ()
        got = data.read_rows(path, names, conf, native=True)
        for i, rec in enumerate(ref):
            mine = got.image(i)
            if len(rec) == 0:
                assert len(mine) == 0
                continue
            assert np.array_equal(rec[0], mine[:, 0].astype(int))
            assert np.array_equal(rec[1], data.xywh_to_xyxy(mine[:, 1:5]))
            if conf:
                assert np.array_equal(rec[2], mine[:, 5])


def test_native_reader_random_number_spellings(tmp_path):
    """Every spelling a detector or a conversion script might emit (%g, %e, %f with few or many digits, repr, integers,
    leading '+', trailing '.', huge and tiny magnitudes) parses to the same float64 bits as Python's float()."""
    rng = np.random.default_rng(12)
    d = tmp_path / "dets"
    d.mkdir()
    fmts = ["%r", "%g", "%.17g", "%e", "%.3e", "%f", "%.10f", "%.1f", "%+g", "%d."]
    names = []
    for i in range(40):
        rows = []
        for _ in range(int(rng.integers(0, 30))):
            vals = np.concatenate([[float(rng.integers(0, 80))], rng.random(4), [rng.random() * 10.0 ** rng.integers(-12, 3)]])
            toks = []
            for v in vals:
                f = fmts[int(rng.integers(len(fmts)))]
                toks.append(repr(float(v)) if f == "%r" else (f % int(v) if f == "%d." else f % v))
            rows.append(" ".join(toks))
        name = f"img{i:03d}"
        names.append(name)
        (d / f"{name}.txt").write_text("\n".join(rows) + ("\n" if rows and rng.random() < 0.7 else ""))
    py = data.read_rows(str(d), names, True, native=False)
    nat = data.read_rows(str(d), names, True, native=True)
    assert same_rows(nat, py) and len(py.rows) > 100
    off, rows, fb = _io.read_rows(str(d), names, True)
    assert len(fb) == 0                                         # all of it on the fast path
