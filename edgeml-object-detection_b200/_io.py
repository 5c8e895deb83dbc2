"""ctypes binding of liborie_io.so (C ABI declared in include/orie_io.h): the native,
multi-threaded reader of the reference's label / detection files."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None
SYMBOLS = ["orie_io_last_error", "orie_io_read_rows", "orie_io_num_images", "orie_io_num_rows", "orie_io_num_cols",
           "orie_io_offsets", "orie_io_data", "orie_io_num_fallback", "orie_io_fallback", "orie_io_free"]


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_build.IO_LIB) or _build.io_stale():
        try:
            _build.build_io()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"liborie_io.so is missing or older than its sources and could not be rebuilt ({e}); "
                               "pass native=False to use the Python reader") from e
    lib = C.CDLL(_build.IO_LIB)
    vp, i64 = C.c_void_p, C.c_int64
    lib.orie_io_last_error.restype = C.c_char_p
    lib.orie_io_read_rows.restype = C.c_int
    lib.orie_io_read_rows.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), i64, C.c_int, C.c_int, C.POINTER(vp)]
    for name, res in (("orie_io_num_images", i64), ("orie_io_num_rows", i64), ("orie_io_num_cols", C.c_int),
                      ("orie_io_offsets", C.POINTER(C.c_int64)), ("orie_io_data", C.POINTER(C.c_double)),
                      ("orie_io_num_fallback", i64), ("orie_io_fallback", C.POINTER(C.c_int64))):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = [vp]
    lib.orie_io_free.restype = None
    lib.orie_io_free.argtypes = [vp]
    _LIB = lib
    return lib


def read_rows(path: str, names, with_conf: bool, threads: int = 0):
    """(off int64[M+1], rows f64[n, 5 or 6], fallback int64[k]) for the images ``names`` of directory ``path``.
    Images listed in ``fallback`` have zero rows here and must be re-read by the caller."""
    lib = load()
    n = len(names)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(s) for s in names])
    h = C.c_void_p(0)
    rc = lib.orie_io_read_rows(os.fsencode(path), arr, n, 1 if with_conf else 0, int(threads), C.byref(h))
    if rc != 0:
        raise OSError(f"orie_io error {rc}: {lib.orie_io_last_error().decode('utf-8', 'replace')}")
    try:
        rows, cols = int(lib.orie_io_num_rows(h)), int(lib.orie_io_num_cols(h))
        off = np.ctypeslib.as_array(lib.orie_io_offsets(h), shape=(n + 1,)).copy()
        data = (np.ctypeslib.as_array(lib.orie_io_data(h), shape=(rows, cols)).copy() if rows
                else np.zeros((0, cols), dtype=np.float64))
        nf = int(lib.orie_io_num_fallback(h))
        fb = np.ctypeslib.as_array(lib.orie_io_fallback(h), shape=(nf,)).copy() if nf else np.zeros(0, dtype=np.int64)
    finally:
        lib.orie_io_free(h)
    return off, data, fb
