"""Seeded synthetic YOLOv5-format datasets (SURVEY.md §8d recipe).

The reference ships no data, so benchmarks and tests run on synthetic
COCO-/VOC-shaped detections.  A dataset is held in the *file-content* form the
reference's loader sees (``lib/data.py:11-43``): per image a block of rows
``cls xc yc w h [conf]`` (normalised xywh, float64), row order = descending
confidence (what YOLOv5 ``val.py --save-txt --save-conf`` and
``torch_models/detect.py:83-105`` write).  Confidences are globally unique and
ground-truth boxes distinct, so the reference's unstable sorts have a single
valid answer and parity is well defined.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Rows:
    """CSR block of per-image rows.  ``off`` int64[M+1]; ``rows`` f64[n, k]."""
    off: np.ndarray
    rows: np.ndarray

    @property
    def num_images(self) -> int:
        return len(self.off) - 1

    def image(self, i: int) -> np.ndarray:
        return self.rows[self.off[i]:self.off[i + 1]]


@dataclass
class SynthDataset:
    names: list
    labels: Rows      # cls xc yc w h
    weak: Rows        # cls xc yc w h conf
    strong: Rows      # cls xc yc w h conf
    num_classes: int
    meta: dict = field(default_factory=dict)


@dataclass
class DetectorShape:
    recall: float
    sigma: float
    mean_dets: float
    max_dets: int
    fixed: bool = False   # exactly max_dets rows per image


# name -> (M, nc, G mean, empty-image fraction, weak, strong, seed)      SURVEY.md §8d table
CONFIGS = {
    "smoke500": (500, 80, 7.4, 0.01, DetectorShape(.55, .08, 60, 300), DetectorShape(.80, .04, 55, 300), 1234),
    "coco5000": (5000, 80, 7.36, 0.01, DetectorShape(.55, .08, 120, 300), DetectorShape(.80, .04, 100, 300), 2017),
    "voc4952": (4952, 20, 2.43, 0.0, DetectorShape(.50, .10, 250, 300), DetectorShape(.85, .03, 15, 100), 2007),
    "sweep50k": (50000, 80, 7.4, 0.01, DetectorShape(1.0, .08, 300, 300, True), DetectorShape(1.0, .04, 300, 300, True), 50000),
}


def _boxes(rng, n):
    c = rng.uniform(0.15, 0.85, size=(n, 2))
    s = np.clip(np.exp(rng.normal(-1.8, 0.7, size=(n, 2))), 0.01, 0.9)
    return np.concatenate([c, s], axis=1)


def _detector(rng, lab_off, lab_rows, nc, shape: DetectorShape):
    M = len(lab_off) - 1
    G = np.diff(lab_off)
    img_of_gt = np.repeat(np.arange(M), G)
    # recalled ground truth -> jittered true-ish detections
    keep = rng.random(len(lab_rows)) < shape.recall
    src = lab_rows[keep]
    img_tp = img_of_gt[keep]
    k = len(src)
    box = src[:, 1:5].copy()
    box[:, 0:2] += rng.normal(0, 1, size=(k, 2)) * shape.sigma * box[:, 2:4]
    box[:, 2:4] *= np.exp(rng.normal(0, shape.sigma, size=(k, 2)))
    box[:, 2:4] = np.clip(box[:, 2:4], 0.005, 0.95)
    cls = src[:, 0].copy()
    flip = rng.random(k) < 0.08
    if nc > 1:
        cls[flip] = (cls[flip] + rng.integers(1, nc, size=int(flip.sum()))) % nc
    conf = rng.beta(3, 2, size=k)
    n_tp = np.bincount(img_tp, minlength=M)
    # false positives
    if shape.fixed:
        n_fp = np.maximum(shape.max_dets - n_tp, 0)
    else:
        lam = np.maximum(shape.mean_dets - shape.recall * G, 0.0)
        n_fp = np.minimum(rng.poisson(lam), np.maximum(shape.max_dets - n_tp, 0))
    f = int(n_fp.sum())
    img_fp = np.repeat(np.arange(M), n_fp)
    fbox = _boxes(rng, f)
    fcls = rng.integers(0, nc, size=f).astype(np.float64)
    fconf = 0.001 + 0.6 * rng.beta(1, 6, size=f)
    img = np.concatenate([img_tp, img_fp])
    rows = np.concatenate([
        np.concatenate([cls[:, None], box, conf[:, None]], axis=1),
        np.concatenate([fcls[:, None], fbox, fconf[:, None]], axis=1)], axis=0)
    # globally unique confidences (nudge exact duplicates)
    c = rows[:, 5]
    while True:
        _, first, cnt = np.unique(c, return_index=True, return_counts=True)
        if len(first) == len(c):
            break
        dup = np.setdiff1d(np.arange(len(c)), first)
        c[dup] = np.nextafter(c[dup], 0.0) - rng.random(len(dup)) * 1e-9
        c[:] = np.clip(c, 1e-6, 1.0)
    # file order: image-major, confidence descending
    order = np.lexsort((-c, img))
    rows, img = rows[order], img[order]
    # cap rows per image (keeps the highest-confidence ones)
    cnt = np.bincount(img, minlength=M)
    start = np.concatenate([[0], np.cumsum(cnt)])
    rank = np.arange(len(img)) - start[img]
    sel = rank < shape.max_dets
    rows, img = rows[sel], img[sel]
    cnt = np.bincount(img, minlength=M)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return Rows(off, np.ascontiguousarray(rows))


def generate(num_images, num_classes, gt_mean, empty_frac, weak: DetectorShape, strong: DetectorShape,
             seed, zipf: float = 0.0, empty_det_frac: float = 0.0) -> SynthDataset:
    """Build a dataset.  ``zipf`` > 0 skews the class prior (rank^-zipf);
    ``empty_det_frac`` blanks that share of detection files per detector."""
    rng = np.random.default_rng(seed)
    M = int(num_images)
    G = np.minimum(rng.geometric(1.0 / max(gt_mean, 1.0), size=M), 80)
    if empty_frac > 0:
        G[rng.random(M) < empty_frac] = 0
    g = int(G.sum())
    if zipf > 0:
        p = np.arange(1, num_classes + 1, dtype=np.float64) ** (-zipf)
        cls = rng.choice(num_classes, size=g, p=p / p.sum())
    else:
        cls = rng.integers(0, num_classes, size=g)
    lab_rows = np.concatenate([cls[:, None].astype(np.float64), _boxes(rng, g)], axis=1)
    lab_off = np.concatenate([[0], np.cumsum(G)]).astype(np.int64)
    labels = Rows(lab_off, lab_rows)
    w = _detector(rng, lab_off, lab_rows, num_classes, weak)
    s = _detector(rng, lab_off, lab_rows, num_classes, strong)
    if empty_det_frac > 0:
        w = _blank(rng, w, empty_det_frac)
        s = _blank(rng, s, empty_det_frac)
    names = [f"{i:012d}" for i in range(M)]
    return SynthDataset(names, labels, w, s, int(num_classes), {"seed": int(seed)})


def _blank(rng, r: Rows, frac: float) -> Rows:
    M = r.num_images
    drop = rng.random(M) < frac
    cnt = np.diff(r.off)
    keep_rows = ~np.repeat(drop, cnt)
    cnt = np.where(drop, 0, cnt)
    return Rows(np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64), np.ascontiguousarray(r.rows[keep_rows]))


def make(config: str, num_images: int | None = None, seed: int | None = None, **kw) -> SynthDataset:
    M, nc, g, e, w, s, sd = CONFIGS[config]
    ds = generate(num_images or M, nc, g, e, w, s, sd if seed is None else seed, **kw)
    ds.meta["config"] = config
    return ds


def _fmt_row(row, with_conf):
    cls = str(int(row[0]))
    return " ".join([cls] + [repr(float(v)) for v in row[1:6 if with_conf else 5]])


def write_dirs(ds: SynthDataset, root: str, strong_as_npy: bool = True):
    """Write ``root/{labels,weak,strong}`` in the reference's on-disk formats:
    ``.txt`` for labels and weak detections (shortest round-trip floats, single
    spaces), ``.npy`` float64[n,6] for strong detections so that both loader
    branches (lib/data.py:23-28) are exercised.  Images without rows get an
    empty label file and *no* detection file."""
    dirs = {k: os.path.join(root, k) for k in ("labels", "weak", "strong")}
    for d in dirs.values():
        os.makedirs(d, exist_ok=True)
    for i, name in enumerate(ds.names):
        lab = ds.labels.image(i)
        with open(os.path.join(dirs["labels"], name + ".txt"), "w") as f:
            f.write("".join(_fmt_row(r, False) + "\n" for r in lab))
        w = ds.weak.image(i)
        if len(w):
            with open(os.path.join(dirs["weak"], name + ".txt"), "w") as f:
                f.write("".join(_fmt_row(r, True) + "\n" for r in w))
        s = ds.strong.image(i)
        if len(s):
            if strong_as_npy:
                np.save(os.path.join(dirs["strong"], name + ".npy"), s)
            else:
                with open(os.path.join(dirs["strong"], name + ".txt"), "w") as f:
                    f.write("".join(_fmt_row(r, True) + "\n" for r in s))
    return dirs["weak"], dirs["strong"], dirs["labels"]
