"""Python call surface mirroring the reference's:

    set_data(weak, strong, label)      lib/data.py:46-84   -> list-of-tuples cache
    compute_rewards_from_dirs(...)     reward.py:72-93     -> reward vector + seconds
    save_rewards(...)                  reward.py:90-92     -> orie{N}.npz / dcsb.npz

Everything numeric runs in the CUDA engine; there is no CPU fallback.
"""
from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np

from . import data
from .engine import (IOU_05, IOU_05_095, Engine, class_shard, clamp_ensemble, combine_sums, pick_shard, rewards_from_sums,
                     shard_of_rank, shard_plan, shard_range)


def parse_iou_thresholds(spec) -> np.ndarray:
    """'0.5' -> [0.5] (what the reference ships, lib/data.py:61);
    '0.5:0.95' -> np.linspace(0.5, 0.95, 10) (the commented alternative, lib/data.py:62);
    'a,b,c' -> explicit list."""
    if isinstance(spec, (list, tuple, np.ndarray)):
        return np.asarray(spec, dtype=np.float64)
    spec = str(spec).strip()
    if spec in ("0.5", ".5"):
        return IOU_05.copy()
    if spec in ("0.5:0.95", ".5:.95", "coco"):
        return IOU_05_095.copy()
    return np.array([float(x) for x in spec.split(",")], dtype=np.float64)


def ensemble_matrix_numpy(num_images: int, num_ensemble: int, base_seed: int, t0: int = 0, nt=None) -> np.ndarray:
    """The reference's draw (reward.py:35-38: arange, shift past the target,
    ``np.random.permutation(...)[:N]``) with the legacy generator seeded by
    ``base_seed + img_idx`` for each target — the parity mode's ensembles."""
    nt = num_images - t0 if nt is None else nt
    n = clamp_ensemble(num_images, num_ensemble)
    out = np.empty((nt, n), dtype=np.int32)
    for r in range(nt):
        i = t0 + r
        others = np.arange(num_images - 1)
        if i < num_images - 1:
            others[i:] += 1
        out[r] = np.random.RandomState((base_seed + i) % 2**32).permutation(others)[:n]
    return out


def set_data(weak, strong, label, iouv=IOU_05, device=None):
    """Drop-in for ``lib.data.set_data`` (same return layout):
    ``weak_data[i] = (tp bool[n,T], conf f64[n], cls int64[n])`` (empty image:
    ``(bool[0,T], f64[0], f64[0])``), same for strong, ``labels[i] = cls
    int64[m]`` or an empty float array.  TP flags come from the CUDA matcher."""
    names, lab, wk, st = data.load_dirs(weak, strong, label)
    pk = data.pack(lab, wk, st)
    iouv = parse_iou_thresholds(iouv)
    eng = Engine(pk, iouv=iouv, device=device)
    try:
        wtp, stp, _, _ = eng.tp_flags()
    finally:
        eng.close()
    T = len(iouv)

    def cache(rows, tp):
        out = []
        for i in range(rows.num_images):
            a, b = rows.off[i], rows.off[i + 1]
            if a == b:
                out.append((np.zeros((0, T), dtype=bool), np.array([]), np.array([])))
            else:
                out.append((tp[a:b], rows.rows[a:b, 5].copy(), rows.rows[a:b, 0].astype(int)))
        return out

    labels = [lab.rows[lab.off[i]:lab.off[i + 1], 0].astype(int) if lab.off[i + 1] > lab.off[i] else np.array([])
              for i in range(lab.num_images)]
    return cache(wk, wtp), cache(st, stp), labels


def compute_rewards_from_dirs(weak_dir, strong_dir, label_dir, method="orie", num_ensemble=1000, iouv=IOU_05,
                              seed=None, ensembles="device", device=None, verbose=True, shard="auto"):
    """What ``reward.py:main`` does between parsing and saving.

    Returns (reward ndarray, seconds, info).  ``seconds`` (saved as ``time``) covers AT LEAST what the
    reference's own timer covers (the reward phase, reward.py:76-88): the dataset sort / index build that
    replaces upstream's per-target argsort, the ensemble draw and the reward kernels — and, because the engine
    overlaps it with the sort, also the upload and TP matching that upstream's timer leaves out (``set_data``).
    File loading is reported separately in ``info`` (``load_s``); ``match_index_s`` is the part of ``seconds``
    spent before the first reward kernel.
    Under torchrun (WORLD_SIZE > 1) the work is split into class groups x target blocks (``engine.shard_plan``:
    classes only up to ~28 k images) and the per-target AP sums are combined with one NCCL all-reduce."""
    import torch
    method = method.lower()
    if method == "ori":
        method, num_ensemble = "orie", 0
    if method not in ("orie", "dcsb"):
        raise ValueError(f"unknown method {method!r}")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group("nccl")
    t = time.perf_counter()
    names, lab, wk, st = data.load_dirs(weak_dir, strong_dir, label_dir)
    pk = data.pack(lab, wk, st)
    t_load = time.perf_counter() - t
    M = pk.num_images
    if M == 0:
        return (np.zeros(0) if method == "orie" else np.zeros(0, dtype=int)), 0.0, {"load_s": t_load}
    iouv = parse_iou_thresholds(iouv)
    N = num_ensemble
    if method == "orie":
        if N > M - 1 and verbose and rank == 0:
            print("Ensemble size is too large. Set to the dataset size.")
        if N < 0 and verbose and rank == 0:
            print("Ensemble size is negative. Set to 0.")
    if seed is None:
        seed = int.from_bytes(os.urandom(4), "little")    # the reference is unseeded
    if dist is not None:
        s = torch.tensor([seed], dtype=torch.int64, device="cuda")
        dist.broadcast(s, 0)
        seed = int(s.item())
    t = time.perf_counter()
    # Multi-GPU: class groups x target blocks (engine.shard_plan); a rank keeps all images but only the rows of its class
    # group, computes the targets of its block, and one all-reduce of the zero-padded per-target AP sums combines the ranks.
    from .engine import HostPacked
    sharded = dist is not None and method == "orie"
    rc, rc_n, t0, nt = shard_of_rank(rank, M, world, shard, pk.num_classes) if sharded else (0, 1, 0, M)
    eng = Engine(HostPacked(class_shard(pk, rc, rc_n) if rc_n > 1 else pk), iouv=iouv, device=device,
                 index=method != "dcsb")
    torch.cuda.synchronize()
    t_match = time.perf_counter() - t
    try:
        start = t            # the timer includes the index build (upstream sorts inside its timer) and the matching
        if method == "dcsb":
            reward = eng.dcsb().astype(int)
        else:
            em = ensemble_matrix_numpy(M, N, seed) if ensembles == "numpy" else None
            if sharded:
                sums = eng.orie_sums_device(N, ens_matrix=None if em is None else em[t0:t0 + nt], seed=seed, t0=t0, nt=nt,
                                            total_images=M)
                eng.stream.synchronize()
                reward = combine_sums(sums, t0, M, eng.T, clamp_ensemble(M, N)).cpu().numpy()
            else:
                reward = eng.orie_device(N, ens_matrix=em, seed=seed).cpu().numpy()
            eng.check_status()
            reward = np.where(np.isnan(reward), 0, reward)     # reward.py:86 (the kernel already stores 0)
        seconds = time.perf_counter() - start
        info = {"load_s": t_load, "match_index_s": t_match, "seed": seed, "images": M, "names": names,
                "num_ensemble_used": clamp_ensemble(M, N) if method == "orie" else None, "index": dict(eng.info)}
    finally:
        eng.close()
    return reward, seconds, info


def rank_normalize(reward, val_mask=None, device=None) -> np.ndarray:
    """Rank-normalised rewards of one cross-validation fold, as ``regression.py:439-441`` computes them when
    ``--normalize`` is set: train rows get ``(argsort(argsort(train)) + 1) / len(train)``, validation rows the share
    of train rewards not above their own.  ``val_mask``: bool[M] (None = every row is a train row).  Runs on the GPU
    (``orie_rank_normalize``); equal train rewards rank in row order."""
    import ctypes as C
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("orie_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = _lib.load()
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    r = torch.from_numpy(np.ascontiguousarray(reward, dtype=np.float64)).to(dev)
    M = int(r.numel())
    out = torch.empty(M, dtype=torch.float64, device=dev)
    mask = None
    if val_mask is not None:
        vm = np.ascontiguousarray(val_mask, dtype=bool)
        if vm.shape != (M,):
            raise ValueError(f"val_mask must have shape ({M},), got {vm.shape}")
        mask = torch.from_numpy(vm.view(np.uint8)).to(dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(lib.orie_rank_workspace_bytes(M)), dtype=torch.uint8, device=dev)
        _lib.check(lib.orie_rank_normalize(C.c_void_p(r.data_ptr()), C.c_void_p(mask.data_ptr() if mask is not None else 0), M,
                                           C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy()


def fit_dcsb(packed, reward, val_mask=None, area_steps=None, n_count: int = 10, device=None) -> dict:
    """The DCSB baseline estimator of one cross-validation fold, as ``baseline.py:67-152`` (``fit_dcsb``) computes it
    from the weak detector's outputs: confidence threshold by bisection, grid search over (object count, smallest
    area) thresholds on the train rows, offloading decisions for every image.

    ``packed``: the ``Packed`` dataset (its weak block and label counts are used); ``reward``: the reward vector
    (binarised ``> 0`` as ``baseline.py:166``); ``val_mask``: bool[M] validation rows of the fold (None = all train).
    Returns ``conf_thresh, num_thresh, area_thresh, train_est, val_est, est`` (``est`` = decisions in image order).
    Runs on the GPU (``orie_dcsb_fit``)."""
    import ctypes as C
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("orie_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = _lib.load()
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    M, D = int(packed.num_images), int(len(packed.w_cls))
    reward = np.asarray(reward)
    if reward.shape != (M,):
        raise ValueError(f"reward must have shape ({M},), got {reward.shape}")
    vm = np.zeros(M, dtype=bool) if val_mask is None else np.ascontiguousarray(val_mask, dtype=bool)
    if vm.shape != (M,):
        raise ValueError(f"val_mask must have shape ({M},), got {vm.shape}")
    steps = np.ascontiguousarray(np.arange(0.2, 0.9, 0.01) if area_steps is None else area_steps, dtype=np.float64)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    box, conf, off = up(packed.w_box), up(packed.w_conf), up(packed.w_off)
    mask = up(vm.view(np.uint8))
    labels = up(np.diff(packed.l_off).astype(np.int64))
    r01 = up(np.where(reward > 0, 1, 0).astype(np.int64))
    model = torch.zeros(4, dtype=torch.float64, device=dev)
    est = torch.zeros(M, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(lib.orie_dcsb_fit_workspace_bytes(M, D)), dtype=torch.uint8, device=dev)
        _lib.check(lib.orie_dcsb_fit(C.c_void_p(box.data_ptr()), C.c_void_p(conf.data_ptr()), C.c_void_p(off.data_ptr()), M, D,
                                     C.c_void_p(mask.data_ptr()), C.c_void_p(labels.data_ptr()), C.c_void_p(r01.data_ptr()),
                                     steps.ctypes.data_as(C.POINTER(C.c_double)), len(steps), int(n_count),
                                     C.c_void_p(model.data_ptr()), C.c_void_p(est.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    m, e = model.cpu().numpy(), est.cpu().numpy()
    return {"conf_thresh": float(m[0]), "num_thresh": int(m[1]), "area_thresh": float(m[2]), "train_hits": int(m[3]),
            "train_est": e[~vm], "val_est": e[vm], "est": e}


def save_rewards(save_dir, method, num_ensemble, reward, seconds):
    """reward.py:90-92: ``orie{N}.npz`` (N as typed on the command line, not the
    clamped value) or ``dcsb.npz`` with keys ``reward`` and ``time``."""
    Path(save_dir).mkdir(parents=True, exist_ok=True)
    method = method.lower()
    if method == "ori":
        method, num_ensemble = "orie", 0
    file_name = f"orie{num_ensemble}.npz" if method == "orie" else "dcsb.npz"
    path = os.path.join(save_dir, file_name)
    np.savez(path, reward=reward, time=seconds)
    return path
