"""Build liborie_b200.so in-tree with nvcc for sm_100a (no other target), and the host-only
file reader liborie_io.so with g++."""
from __future__ import annotations

import contextlib
import fcntl
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liborie_b200.so")
SOURCES = ["api.cu", "sort.cu", "match.cu", "index.cu", "reward.cu", "rank.cu", "dcsb_fit.cu"]
HEADERS = ["common.cuh", "coop.cuh", "index.cuh", os.path.join("..", "..", "include", "orie_b200.h")]
IO_LIB = os.path.join(HERE, "liborie_io.so")
IO_SOURCES = ["loader.cpp"]
IO_HEADERS = [os.path.join("..", "..", "include", "orie_io.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                    # FP64 parity: never contract a*b+c
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared", "-cudart", "static",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the engine has no CPU fallback and cannot be built without it")
    return exe


def nvcc_or_none():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return exe if os.path.exists(exe) else None


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


@contextlib.contextmanager
def _build_lock():
    """One builder at a time per checkout (every rank of a torchrun job imports this module at once)."""
    with open(os.path.join(HERE, ".build.lock"), "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def _compile(cmd, out) -> str:
    """Run ``cmd + ['-o', tmp]`` and move the result into place atomically, so a process that is loading
    ``out`` never sees a half-written library."""
    tmp = f"{out}.{os.getpid()}.tmp"
    try:
        res = subprocess.run(cmd + ["-o", tmp], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    with _build_lock():
        if not force and not stale():      # another process built it while this one waited for the lock
            return LIB
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, f) for f in SOURCES]
        log = _compile(cmd, LIB)
    if verbose:
        print(log)
    return LIB


def io_stale() -> bool:
    if not os.path.exists(IO_LIB):
        return True
    t = os.path.getmtime(IO_LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in IO_SOURCES + IO_HEADERS)


def build_io(force: bool = False) -> str:
    """liborie_io.so: the native reader of the label / detection files (plain C++17 + pthreads, no CUDA)."""
    if not force and not io_stale():
        return IO_LIB
    gxx = shutil.which("g++") or shutil.which("c++")
    if gxx is None:
        raise RuntimeError("g++ not found; liborie_io.so cannot be built")
    with _build_lock():
        if not force and not io_stale():
            return IO_LIB
        _compile([gxx, "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-pthread"] +
                 [os.path.join(CSRC, f) for f in IO_SOURCES], IO_LIB)
    return IO_LIB


if __name__ == "__main__":
    print(build_io(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
