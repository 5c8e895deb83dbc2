"""Realised mAP versus offloading ratio (the reference's ``test.py``), on the same CUDA kernels.

``test.py:39-42`` builds, for every offloading ratio, the detection set "strong output for offloaded images, weak
output for the others" and calls ``ap_per_class`` over the WHOLE dataset.  That is an ORIE evaluation in disguise:
give every image two pseudo-images (2i = its weak detections, 2i+1 = its strong detections, both carrying the image's
labels), let an *empty* pseudo-image be the target and let the ensemble be {2i + offload_i}; then
``mAP(E_weak ∪ weak_target) = mAP(E)`` is exactly the dataset mAP of that selection, and exactly one copy of every
label is counted.  So this module only rearranges rows on the host and calls the engine (``orie_match``,
``orie_index_build``, ``orie_ensemble_from_indices``, ``orie_reward`` with its per-target AP sums) — no extra kernels.
"""
from __future__ import annotations

import os

import numpy as np

from . import data
from .engine import IOU_05, Engine
from .synth import Rows

OFFLOADING_RATIOS = np.arange(0, 1.01, 0.1)          # test.py:11


def _gather_blocks(src_rows, starts, lens):
    """Concatenate src_rows[starts[k] : starts[k] + lens[k]] for all k."""
    total = int(lens.sum())
    if total == 0:
        return np.zeros((0, src_rows.shape[1]))
    out_off = np.concatenate([[0], np.cumsum(lens)])[:-1]
    idx = np.repeat(starts - out_off, lens) + np.arange(total)
    return src_rows[idx]


def pseudo_dataset(labels: Rows, weak: Rows, strong: Rows, num_targets: int):
    """Rows of the pseudo-dataset and the index of its first (empty) target image."""
    M = labels.num_images
    Dw = len(weak.rows)
    stacked = np.concatenate([weak.rows, strong.rows], axis=0) if Dw + len(strong.rows) else np.zeros((0, 6))
    cw, cs, cl = np.diff(weak.off), np.diff(strong.off), np.diff(labels.off)
    t0 = (2 * M + 31) // 32 * 32                       # targets start on a multiple of 32
    Mp = t0 + num_targets
    det_counts = np.zeros(Mp, dtype=np.int64)
    det_counts[0:2 * M:2], det_counts[1:2 * M:2] = cw, cs
    starts = np.zeros(2 * M, dtype=np.int64)
    starts[0::2], starts[1::2] = weak.off[:-1], Dw + strong.off[:-1]
    dets = Rows(np.concatenate([[0], np.cumsum(det_counts)]).astype(np.int64),
                _gather_blocks(stacked, starts, det_counts[:2 * M]))
    lab_counts = np.zeros(Mp, dtype=np.int64)
    lab_counts[0:2 * M:2] = lab_counts[1:2 * M:2] = cl
    lab = Rows(np.concatenate([[0], np.cumsum(lab_counts)]).astype(np.int64),
               _gather_blocks(labels.rows, np.repeat(labels.off[:-1], 2), lab_counts[:2 * M]))
    none = Rows(np.zeros(Mp + 1, dtype=np.int64), np.zeros((0, 6)))
    return lab, dets, none, t0


def realised_map(labels: Rows, weak: Rows, strong: Rows, offload_masks, iouv=IOU_05, device=None) -> np.ndarray:
    """mAP of the dataset for every row of ``offload_masks`` (bool[R, M]; True = the image uses the strong
    detector's output) == ``np.mean(ap_per_class(...))`` of test.py:39-42."""
    masks = np.atleast_2d(np.asarray(offload_masks, dtype=bool))
    R, M = masks.shape
    if M != labels.num_images:
        raise ValueError("offload mask length differs from the number of images")
    lab, dets, none, t0 = pseudo_dataset(labels, weak, strong, R)
    pk = data.pack(lab, dets, none)
    eng = Engine(pk, iouv=iouv, device=device)
    try:
        ens = (2 * np.arange(M, dtype=np.int32)[None, :] + masks.astype(np.int32))
        _, detail = eng.orie(M, ens_matrix=ens, t0=t0, nt=R, detail=True)
        T = eng.T
    finally:
        eng.close()
    with np.errstate(invalid="ignore", divide="ignore"):
        return detail[:, 0] / (detail[:, 2] * T)        # mean over the (classes with labels) x T table; NaN if no labels


def offload_masks_from_estimates(estimate_dir: str, dataset_split: np.ndarray, ratios=OFFLOADING_RATIOS) -> np.ndarray:
    """bool[len(ratios), M]: the fixed-threshold policy of test.py:26-37 — per cross-validation fold the threshold
    is the estimated reward of the training image at rank int((n_train - 1) * ratio) (descending), and a
    validation image is offloaded iff its estimate exceeds it."""
    dataset_split = np.asarray(dataset_split, dtype=bool)
    masks = np.zeros((len(ratios), dataset_split.shape[1]), dtype=bool)
    for fold, val_mask in enumerate(dataset_split):
        est = np.load(os.path.join(estimate_dir, f"estimate{fold + 1}.npz"))
        train, val = est["train_est"], est["val_est"]
        ranked = train[np.argsort(-train)]
        for r, ratio in enumerate(ratios):
            masks[r, val_mask] = val > ranked[int((len(train) - 1) * ratio)]
    return masks


def test_map_from_dirs(weak_dir, strong_dir, label_dir, split_path, estimates, iouv=IOU_05, device=None) -> np.ndarray:
    """f64[len(estimates), 11] — what the reference's test.py saves as test_map.npy."""
    _, lab, wk, st = data.load_dirs(weak_dir, strong_dir, label_dir)
    split = np.load(split_path)
    estimates = [estimates] if isinstance(estimates, str) else list(estimates or [])
    if not estimates:
        return np.zeros((0, len(OFFLOADING_RATIOS)))
    masks = np.concatenate([offload_masks_from_estimates(e, split) for e in estimates], axis=0)
    maps = realised_map(lab, wk, st, masks, iouv=iouv, device=device)
    return maps.reshape(len(estimates), len(OFFLOADING_RATIOS))
