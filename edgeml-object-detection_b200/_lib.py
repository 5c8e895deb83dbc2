"""ctypes binding of liborie_b200.so (C ABI declared in include/orie_b200.h).

There is no CPU fallback: if the shared library is missing and cannot be built
with nvcc, importing the engine fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_i64p = C.c_void_p   # device pointers are passed as plain addresses
SYMBOLS = [
    "orie_last_error", "orie_version", "orie_match", "orie_dcsb",
    "orie_index_build", "orie_index_sizes", "orie_index_build_into", "orie_index_destroy", "orie_index_set_aux_stream", "orie_index_info", "orie_index_status",
    "orie_ensemble_from_indices", "orie_ensemble_sample", "orie_ensemble_sample_dev",
    "orie_reward_workspace_bytes", "orie_reward_workspace_bound", "orie_reward", "orie_reward_sums", "orie_rewards_from_sums",
    "orie_reward_profile", "orie_reward_depths", "orie_launch_count",
    "orie_rank_workspace_bytes", "orie_rank_normalize", "orie_dcsb_fit_workspace_bytes", "orie_dcsb_fit",
]


class IndexInfo(C.Structure):
    _fields_ = [
        ("num_images", C.c_int64), ("num_classes", C.c_int64),
        ("num_thresholds", C.c_int32), ("seg_chunks", C.c_int32),
        ("num_weak", C.c_int64), ("num_strong", C.c_int64), ("num_labels", C.c_int64),
        ("slots", C.c_int64), ("segments", C.c_int64), ("events", C.c_int64),
        ("label_slots", C.c_int64), ("label_segments", C.c_int64),
        ("class_groups", C.c_int64), ("ens_words", C.c_int64), ("device_bytes", C.c_int64),
    ]


class Tuning(C.Structure):
    _fields_ = [("seg_chunks", C.c_int32), ("sort_max_blocks", C.c_int32), ("post_blocks", C.c_int32),
                ("walk_gmem", C.c_int32), ("ap_mode", C.c_int32), ("sort_lsd", C.c_int32), ("walk_unpacked", C.c_int32), ("walk_single", C.c_int32),
                ("walk_waves", C.c_double)]


class OrieError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"orie_b200 error {code}: {message}")
        self.code = code


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources are newer) and type the library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path) or (_build.stale() and os.path.exists(_build.nvcc_or_none() or "")):
        # sources newer than the library (or no library): rebuild under the build lock.  A failed rebuild is an
        # error even if an older library is still lying around — never run a binary that is not the sources'.
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(
                "liborie_b200.so is missing or older than its sources and could not be rebuilt; the engine has no "
                f"CPU fallback. Build it with `python -c 'import __graft_entry__ as g; g.build()'` ({e})") from e
    lib = C.CDLL(path)
    vp, i64, i32, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64
    lib.orie_last_error.restype = C.c_char_p
    lib.orie_last_error.argtypes = []
    lib.orie_version.restype = C.c_int
    lib.orie_match.restype = C.c_int
    lib.orie_match.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(C.c_double), i32, i64, vp, vp, vp, vp]
    lib.orie_dcsb.restype = C.c_int
    lib.orie_dcsb.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    lib.orie_index_build.restype = C.c_int
    lib.orie_index_build.argtypes = [i64, i64, i32, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(Tuning), vp, vp,
                                     C.POINTER(vp)]
    lib.orie_index_sizes.restype = C.c_int
    lib.orie_index_sizes.argtypes = [i64, i64, i32, i64, i64, i64, C.POINTER(Tuning), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.orie_index_build_into.restype = C.c_int
    lib.orie_index_build_into.argtypes = [i64, i64, i32, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(Tuning),
                                          vp, C.c_size_t, vp, C.c_size_t, vp, vp, C.POINTER(vp)]
    lib.orie_ensemble_sample_dev.restype = C.c_int
    lib.orie_ensemble_sample_dev.argtypes = [vp, i64, i64, i64, vp, vp, vp]
    lib.orie_index_set_aux_stream.restype = C.c_int
    lib.orie_index_set_aux_stream.argtypes = [vp, vp]
    lib.orie_index_status.restype = C.c_int
    lib.orie_index_status.argtypes = [vp]
    lib.orie_reward_workspace_bound.restype = C.c_size_t
    lib.orie_reward_workspace_bound.argtypes = [vp, i64]
    lib.orie_reward_depths.restype = C.c_int
    lib.orie_reward_depths.argtypes = [vp, i64, i64, vp, i64, vp, C.c_size_t, vp, vp, vp]
    lib.orie_rewards_from_sums.restype = C.c_int
    lib.orie_rewards_from_sums.argtypes = [vp, i64, i32, i64, vp, vp]
    lib.orie_index_destroy.restype = None
    lib.orie_index_destroy.argtypes = [vp]
    lib.orie_index_info.restype = C.c_int
    lib.orie_index_info.argtypes = [vp, C.POINTER(IndexInfo)]
    lib.orie_ensemble_from_indices.restype = C.c_int
    lib.orie_ensemble_from_indices.argtypes = [vp, i64, i64, vp, i64, vp, vp, vp]
    lib.orie_ensemble_sample.restype = C.c_int
    lib.orie_ensemble_sample.argtypes = [vp, i64, i64, i64, u64, vp, vp]
    lib.orie_reward_workspace_bytes.restype = C.c_size_t
    lib.orie_reward_workspace_bytes.argtypes = [vp, i64]
    lib.orie_reward.restype = C.c_int
    lib.orie_reward.argtypes = [vp, i64, i64, vp, i64, vp, C.c_size_t, vp, vp, vp]
    lib.orie_reward_profile.restype = C.c_int
    lib.orie_reward_sums.restype = C.c_int
    lib.orie_reward_sums.argtypes = [vp, i64, i64, vp, i64, vp, C.c_size_t, vp, i32, vp]
    lib.orie_reward_profile.argtypes = [vp, i64, i64, vp, i64, vp, C.c_size_t, vp, vp, i32, vp, C.POINTER(C.c_float)]
    lib.orie_rank_workspace_bytes.restype = C.c_size_t
    lib.orie_rank_workspace_bytes.argtypes = [i64]
    lib.orie_rank_normalize.restype = C.c_int
    lib.orie_rank_normalize.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
    lib.orie_dcsb_fit_workspace_bytes.restype = C.c_size_t
    lib.orie_dcsb_fit_workspace_bytes.argtypes = [i64, i64]
    lib.orie_dcsb_fit.restype = C.c_int
    lib.orie_dcsb_fit.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp, C.POINTER(C.c_double), i32, i32, vp, vp, vp, C.c_size_t, vp]
    lib.orie_launch_count.restype = C.c_longlong
    lib.orie_launch_count.argtypes = []
    _LIB = lib
    return lib


def check(code: int):
    if code != 0:
        raise OrieError(code, load().orie_last_error().decode("utf-8", "replace"))
