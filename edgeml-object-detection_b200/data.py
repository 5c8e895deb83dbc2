"""YOLOv5-format label / detection loaders and the host-side packer.

Mirrors the reference's loader contract (``lib/data.py:11-59``):

* the *label directory listing* defines the image set and its order
  (``sorted(os.listdir(label_dir))`` with the last extension stripped,
  ``lib/data.py:54-56``);
* per image ``<name>.txt`` is read if it exists, else ``<name>.npy``, else the
  image has no rows (``lib/data.py:23-28,39-41``);
* text rows are split on single spaces and parsed as float64; column 0 is the
  class (truncated to int), columns 1..4 are normalised ``xc yc w h`` and, for
  detections, the LAST column is the confidence (``lib/data.py:24-38``);
* malformed numbers raise ``ValueError`` exactly like ``astype(float)`` does;
* row order inside a file is significant (matching tie-break = lower row).

The packer turns the three row blocks into the flat CSR arrays the CUDA
engine consumes: xyxy float64 boxes (``lib/metrics.py:6-18`` arithmetic, in
numpy float64 on the host so the bits equal the reference's), float64
confidences and dense int32 class ids.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .synth import Rows


def list_images(label_dir: str) -> list:
    """Image names in the reference's order (``lib/data.py:54-56``)."""
    return ['.'.join(n.split('.')[:-1]) for n in sorted(os.listdir(label_dir))]


def _parse_text(path: str) -> np.ndarray:
    with open(path, "r") as f:
        lines = f.readlines()
    if not lines:
        return np.zeros((0, 0))
    table = [ln.strip().split(' ') for ln in lines]
    width = len(table[0])
    # zip(*rows) upstream silently truncates to the shortest row; keep that.
    width = min(len(r) for r in table)
    return np.array([r[:width] for r in table]).astype(float).reshape(len(table), width)


def _read_image(stem: str, need: int, with_conf: bool):
    """Rows of one image, the way ``lib/data.py:23-38`` reads them (``.txt`` first, else ``.npy``, else none)."""
    arr = None
    if os.path.isfile(stem + ".txt"):
        arr = _parse_text(stem + ".txt")
    elif os.path.isfile(stem + ".npy"):
        arr = np.asarray(np.load(stem + ".npy"), dtype=float)
        if arr.ndim != 2:
            arr = arr.reshape(len(arr), -1) if arr.size else np.zeros((0, 0))
    if arr is None or len(arr) == 0:
        return None
    if arr.shape[1] < need:
        raise ValueError(f"{stem}: expected at least {need} columns per row, found {arr.shape[1]}")
    if with_conf:
        return np.concatenate([arr[:, 0:5], arr[:, -1:]], axis=1)
    return arr[:, 0:5]


def read_rows_python(path: str, names, with_conf: bool) -> Rows:
    """Pure-Python reader: the reference-equivalent statement of ``lib/data.py:11-43`` (one file at a time)."""
    need = 6 if with_conf else 5
    blocks, counts = [], []
    for name in names:
        arr = _read_image(os.path.join(path, name), need, with_conf)
        if arr is None:
            counts.append(0)
            continue
        blocks.append(arr)
        counts.append(len(arr))
    rows = np.concatenate(blocks, axis=0) if blocks else np.zeros((0, need))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return Rows(off, np.ascontiguousarray(rows, dtype=np.float64))


def read_rows(path: str, names, with_conf: bool, native: bool = True, threads: int = 0) -> Rows:
    """Rows of every image in ``names`` from directory ``path`` as one CSR
    block: labels -> ``cls xc yc w h``; detections -> ``cls xc yc w h conf``.

    ``native=True`` (default) reads with the multi-threaded C++ reader (``csrc/loader.cpp`` through
    ``include/orie_io.h``); the few files it hands back (anything that is not plain numeric text or a little-endian
    f4/f8 ``.npy`` matrix) are re-read by the Python path, which raises what the reference raises."""
    if not native:
        return read_rows_python(path, names, with_conf)
    from . import _io
    need = 6 if with_conf else 5
    off, rows, fallback = _io.read_rows(path, list(names), with_conf, threads)
    if len(fallback):
        extra = {int(i): _read_image(os.path.join(path, names[int(i)]), need, with_conf) for i in fallback}
        extra = {i: a for i, a in extra.items() if a is not None}
        if extra:
            counts = np.diff(off)
            blocks, prev = [], 0
            for i in sorted(extra):
                blocks.append(rows[off[prev]:off[i]])
                blocks.append(np.ascontiguousarray(extra[i], dtype=np.float64))
                counts[i] = len(extra[i])
                prev = i
            blocks.append(rows[off[prev]:])
            rows = np.concatenate(blocks, axis=0)
            off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return Rows(off, np.ascontiguousarray(rows, dtype=np.float64))


def load_dirs(weak_dir: str, strong_dir: str, label_dir: str):
    """(names, labels, weak, strong) row blocks for the three directories."""
    names = list_images(label_dir)
    return names, read_rows(label_dir, names, False), read_rows(weak_dir, names, True), read_rows(strong_dir, names, True)


def xywh_to_xyxy(b: np.ndarray) -> np.ndarray:
    """``lib/metrics.py:6-18`` in float64: x -/+ w/2, y -/+ h/2."""
    out = np.empty((len(b), 4), dtype=np.float64)
    out[:, 0] = b[:, 0] - b[:, 2] / 2
    out[:, 1] = b[:, 1] - b[:, 3] / 2
    out[:, 2] = b[:, 0] + b[:, 2] / 2
    out[:, 3] = b[:, 1] + b[:, 3] / 2
    return out


@dataclass
class Packed:
    """Flat host arrays handed to the engine (all C-contiguous)."""
    num_images: int
    num_classes: int
    class_values: np.ndarray     # int64[C]: dense id -> class value in the files
    w_off: np.ndarray            # int64[M+1]
    w_box: np.ndarray            # f64[Dw,4] xyxy
    w_conf: np.ndarray           # f64[Dw]
    w_cls: np.ndarray            # int32[Dw] dense
    s_off: np.ndarray
    s_box: np.ndarray
    s_conf: np.ndarray
    s_cls: np.ndarray
    l_off: np.ndarray            # int64[M+1]
    l_box: np.ndarray            # f64[G,4] xyxy
    l_cls: np.ndarray            # int32[G] dense

    def nbytes(self) -> int:
        return sum(getattr(self, k).nbytes for k in
                   ("w_off", "w_box", "w_conf", "w_cls", "s_off", "s_box", "s_conf", "s_cls", "l_off", "l_box", "l_cls"))


def pack(labels: Rows, weak: Rows, strong: Rows) -> Packed:
    M = labels.num_images
    if weak.num_images != M or strong.num_images != M:
        raise ValueError("labels / weak / strong blocks cover different image counts")
    lc = labels.rows[:, 0].astype(np.int64)
    wc = weak.rows[:, 0].astype(np.int64)
    sc = strong.rows[:, 0].astype(np.int64)
    values = np.unique(np.concatenate([lc, wc, sc]))
    if len(values) == 0:
        values = np.zeros(1, dtype=np.int64)

    lut = None
    if values[0] >= 0 and values[-1] < (1 << 20):        # the usual case: small non-negative class ids
        lut = np.zeros(int(values[-1]) + 1, dtype=np.int32)
        lut[values] = np.arange(len(values), dtype=np.int32)

    def dense(c):
        return lut[c] if lut is not None else np.searchsorted(values, c).astype(np.int32)

    c64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    return Packed(
        num_images=M, num_classes=int(len(values)), class_values=values,
        w_off=weak.off.astype(np.int64), w_box=xywh_to_xyxy(weak.rows[:, 1:5]), w_conf=c64(weak.rows[:, 5]), w_cls=dense(wc),
        s_off=strong.off.astype(np.int64), s_box=xywh_to_xyxy(strong.rows[:, 1:5]), s_conf=c64(strong.rows[:, 5]), s_cls=dense(sc),
        l_off=labels.off.astype(np.int64), l_box=xywh_to_xyxy(labels.rows[:, 1:5]), l_cls=dense(lc))
