"""CPU oracle for the ORIE / ORI / DCSB offloading-reward path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the
checker or as the timed CPU baseline.

This is a numpy restatement (float64 / int64, the reference's operation order)
of the reference's algorithm.  Each function cites the reference lines it
follows (paths relative to the upstream repository root):

  xywh_to_xyxy      lib/metrics.py:6-18     (xywh2xyxy)
  pairwise_iou      lib/metrics.py:67-86    (box_iou)
  match_detections  lib/metrics.py:38-64    (box_correct)
  ap_101            lib/metrics.py:127-148  (compute_ap, 'interp' branch)
  ap_by_class       lib/metrics.py:89-124   (ap_per_class)
  build_cache       lib/data.py:46-84       (set_data, after loading)
  ensemble_indices  reward.py:28-38         (ensemble draw inside compute_orie)
  orie_one          reward.py:16-52         (compute_orie)
  dcsb_one          reward.py:55-69         (compute_dcsb)

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned by running the *live* reference in the build container
(``oracle/ref_harness.py``; ``tests/test_oracle_vs_reference.py``) and by the
fixtures frozen from those runs under ``tests/golden/`` (generator:
``oracle/gen_golden.py``).

Third-party arithmetic the reference delegates to numpy (un-pinned upstream;
numpy 2.3.5 here): argsort, unique, cumsum, interp, trapz, mean, permutation.
Tie semantics (equal confidences inside one class, equal IoU of one detection
to two same-class labels) are unspecified upstream because ``argsort`` is
unstable; this oracle, like the engine, uses the documented deterministic rule
(stable by concatenation order; IoU ties -> highest label index).
"""
from __future__ import annotations

import warnings

import numpy as np

IOU_05 = np.array([0.5])
IOU_05_095 = np.linspace(0.5, 0.95, 10)
_GRID = np.linspace(0, 1, 101)

_trapz = getattr(np, "trapz", None) or np.trapezoid


def xywh_to_xyxy(b: np.ndarray) -> np.ndarray:
    """(xc, yc, w, h) -> (x1, y1, x2, y2); float64, half-extent = w / 2."""
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
    out = np.empty_like(b)
    hw = b[:, 2] / 2
    hh = b[:, 3] / 2
    out[:, 0] = b[:, 0] - hw
    out[:, 1] = b[:, 1] - hh
    out[:, 2] = b[:, 0] + hw
    out[:, 3] = b[:, 1] + hh
    return out


def pairwise_iou(lab: np.ndarray, det: np.ndarray) -> np.ndarray:
    """IoU[m, n] of label boxes against detection boxes (both xyxy float64).

    Operation order as upstream: inter = max(0, dx) * max(0, dy);
    union = (area_lab[:, None] + area_det) - inter; no epsilon.
    """
    lab = np.asarray(lab, dtype=np.float64).reshape(-1, 4)
    det = np.asarray(det, dtype=np.float64).reshape(-1, 4)
    ix1 = np.maximum(lab[:, None, 0], det[None, :, 0])
    iy1 = np.maximum(lab[:, None, 1], det[None, :, 1])
    ix2 = np.minimum(lab[:, None, 2], det[None, :, 2])
    iy2 = np.minimum(lab[:, None, 3], det[None, :, 3])
    inter = np.maximum(0, ix2 - ix1) * np.maximum(0, iy2 - iy1)
    a_lab = (lab[:, 2] - lab[:, 0]) * (lab[:, 3] - lab[:, 1])
    a_det = (det[:, 2] - det[:, 0]) * (det[:, 3] - det[:, 1])
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / (a_lab[:, None] + a_det[None, :] - inter)


def match_detections(det_box, det_cls, lab_box, lab_cls, iouv):
    """TP flags of one image's detections at every IoU threshold.

    Data-parallel restatement of upstream's sort/unique/unique procedure:
      best[d]  = same-class label with the largest IoU (threshold independent;
                 exact IoU ties -> highest label index),
      d is a TP at threshold t  iff  IoU[best[d], d] >= t  and no detection
      d' < d (file order) has best[d'] == best[d] with IoU >= t.
    A loser never falls back to its second-best label.

    Returns (tp bool[n, T], best_label int32[n] (-1: no same-class label with a
    comparable IoU), best_iou float64[n] (0 when best_label is -1)).
    """
    iouv = np.asarray(iouv, dtype=np.float64)
    n, m, T = len(det_cls), len(lab_cls), len(iouv)
    tp = np.zeros((n, T), dtype=bool)
    best = np.full(n, -1, dtype=np.int32)
    biou = np.zeros(n, dtype=np.float64)
    if n == 0 or m == 0:
        return tp, best, biou
    iou = pairwise_iou(lab_box, det_box)
    same = np.asarray(lab_cls).reshape(-1, 1) == np.asarray(det_cls).reshape(1, -1)
    for d in range(n):
        b, bv = -1, -1.0
        for l in range(m):
            v = iou[l, d]
            if same[l, d] and v >= bv:  # NaN never passes; ties -> later label
                b, bv = l, v
        if b >= 0:
            best[d], biou[d] = b, bv
    for t in range(T):
        taken = np.zeros(m, dtype=bool)
        for d in range(n):
            b = best[d]
            if b >= 0 and biou[d] >= iouv[t] and not taken[b]:
                taken[b] = True
                tp[d, t] = True
    return tp, best, biou


def match_detections_sortunique(det_box, det_cls, lab_box, lab_cls, iouv):
    """Same result via the upstream procedure (stable sort, two 'first of
    each' selections).  Used to cross-check ``match_detections``.
    Returns (tp bool[n, T], per-threshold list of (label, det) pairs)."""
    iouv = np.asarray(iouv, dtype=np.float64)
    n, T = len(det_cls), len(iouv)
    tp = np.zeros((n, T), dtype=bool)
    pairs = []
    if n == 0 or len(lab_cls) == 0:
        return tp, [np.zeros((0, 2), dtype=np.int64) for _ in range(T)]
    iou = pairwise_iou(lab_box, det_box)
    same = np.asarray(lab_cls).reshape(-1, 1) == np.asarray(det_cls).reshape(1, -1)
    for t in range(T):
        li, di = np.nonzero((iou >= iouv[t]) & same)
        if li.size:
            v = iou[li, di]
            order = np.argsort(v, kind="stable")[::-1]
            li, di = li[order], di[order]
            _, first = np.unique(di, return_index=True)
            li, di = li[first], di[first]
            _, first = np.unique(li, return_index=True)
            li, di = li[first], di[first]
            tp[di, t] = True
        pairs.append(np.stack([li, di], axis=1).astype(np.int64))
    return tp, pairs


def ap_101(recall: np.ndarray, precision: np.ndarray) -> float:
    """101-point interpolated AP of one precision/recall curve."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        return float(_trapz(np.interp(_GRID, mrec, mpre), _GRID))


def ap_by_class(tp, conf, pred_cls, target_cls):
    """AP[nc, T] over the classes present in ``target_cls``.

    Follows upstream's structure (global confidence sort, then one masked
    cumulative pass per ground-truth class) so that it is also a fair CPU
    timing stand-in.  Stable sort (documented tie rule)."""
    tp = np.asarray(tp)
    order = np.argsort(-np.asarray(conf, dtype=np.float64), kind="stable")
    tp, pred_cls = tp[order], np.asarray(pred_cls)[order]
    classes, counts = np.unique(target_cls, return_counts=True)
    T = tp.shape[1]
    ap = np.zeros((classes.shape[0], T))
    for row, c in enumerate(classes):
        sel = pred_cls == c
        if not sel.any() or counts[row] == 0:
            continue
        hits = tp[sel]
        tpc = hits.cumsum(0)
        fpc = (1 - hits).cumsum(0)
        recall = tpc / (counts[row] + 1e-16)
        precision = tpc / (tpc + fpc)
        for t in range(T):
            ap[row, t] = ap_101(recall[:, t], precision[:, t])
    return ap


def build_cache(weak, strong, labels, iouv=IOU_05):
    """Per-image cached statistics from loaded (already xyxy) records.

    ``weak`` / ``strong``: lists of (cls int64[n], xyxy f64[n,4], conf f64[n])
    or () ; ``labels``: list of (cls int64[m], xyxy f64[m,4]) or ().
    Returns (weak_data, strong_data, label_cls) in upstream's layout:
    weak_data[i] = (tp bool[n,T], conf f64[n], cls[n]); label_cls[i] = cls[m].
    """
    iouv = np.asarray(iouv, dtype=np.float64)
    T = len(iouv)

    def one(det, lab):
        if len(det) == 0:
            return np.zeros((0, T), dtype=bool), np.array([]), np.array([])
        cls, box, conf = det
        if len(lab) == 0:
            return np.zeros((len(cls), T), dtype=bool), conf, cls
        tp, _, _ = match_detections(box, cls, lab[1], lab[0], iouv)
        return tp, conf, cls

    wd, sd, lc = [], [], []
    for w, s, l in zip(weak, strong, labels):
        wd.append(one(w, l))
        sd.append(one(s, l))
        lc.append(l[0] if len(l) > 0 else np.array([]))
    return wd, sd, lc


def clamp_ensemble(num_img: int, num_ensemble: int) -> int:
    return max(0, min(int(num_ensemble), num_img - 1))


def ensemble_indices(num_img: int, img_idx: int, num_ensemble: int, seed: int) -> np.ndarray:
    """The ensemble upstream would draw for ``img_idx`` if the legacy global
    RNG had been seeded with ``seed`` immediately before the call."""
    n = clamp_ensemble(num_img, num_ensemble)
    others = np.arange(num_img - 1)
    if img_idx < num_img - 1:
        others[img_idx:] += 1
    state = np.random.RandomState(seed % 2**32)
    return state.permutation(others)[:n]


def ensemble_matrix(num_img: int, num_ensemble: int, base_seed: int, targets=None) -> np.ndarray:
    """int32[len(targets), N] explicit ensembles, seed = base_seed + img_idx."""
    targets = range(num_img) if targets is None else targets
    n = clamp_ensemble(num_img, num_ensemble)
    out = np.empty((len(targets), n), dtype=np.int32)
    for r, i in enumerate(targets):
        out[r] = ensemble_indices(num_img, int(i), num_ensemble, base_seed + int(i))
    return out


def orie_one(img_idx, weak_data, strong_data, label_cls, ens_idx):
    """(N+1) * (mAP with the target offloaded - mAP with it kept local).

    Returns (orie, weak_ap[nc,T], strong_ap[nc,T]); orie is NaN when the
    evaluated set has no ground truth (upstream turns NaN into 0 afterwards).
    """
    ens_idx = np.asarray(ens_idx, dtype=np.int64)
    members = list(ens_idx) + [img_idx]
    gt = np.concatenate([label_cls[s] for s in members]).astype(int)

    def gather(last):
        recs = [weak_data[s] for s in ens_idx] + [last]
        return [np.concatenate(col, axis=0) for col in zip(*recs)]

    weak_ap = ap_by_class(*gather(weak_data[img_idx]), gt)
    strong_ap = ap_by_class(*gather(strong_data[img_idx]), gt)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        orie = (np.mean(strong_ap) - np.mean(weak_ap)) * (len(ens_idx) + 1)
    return float(orie), weak_ap, strong_ap


def orie_all(weak_data, strong_data, label_cls, ens_matrix, targets=None):
    """Rewards for ``targets`` (default all) with NaN -> 0 (reward.py:86)."""
    targets = range(len(label_cls)) if targets is None else targets
    out = np.empty(len(targets), dtype=np.float64)
    for r, i in enumerate(targets):
        out[r] = orie_one(int(i), weak_data, strong_data, label_cls, ens_matrix[r])[0]
    return np.where(np.isnan(out), 0, out)


def dcsb_one(img_idx, weak_data, strong_data) -> int:
    return int(np.sum(strong_data[img_idx][1] > 0.5)) - int(np.sum(weak_data[img_idx][1] > 0.5))


def dcsb_all(weak_data, strong_data) -> np.ndarray:
    return np.array([dcsb_one(i, weak_data, strong_data) for i in range(len(weak_data))], dtype=np.int64)
