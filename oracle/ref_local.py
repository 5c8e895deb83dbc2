"""Loader of the staged reference (oracle/_ref/, made by oracle/make_ref.py from /root/reference).

TEST INFRASTRUCTURE ONLY: used by ``bench.py``'s ``--impl reference`` / ``cpu_baseline`` legs to time the reference's
own ``compute_orie`` (reward.py:16-52) on the host cores of the GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import warnings

import numpy as np

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_MODS = None


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in ("reward.py", "lib/metrics.py", "lib/data.py"))


def modules():
    """(reward, lib.metrics) of the staged reference.  ``reward.py`` is loaded by path under another name because this
    repository has a ``reward.py`` of its own (the drop-in CLI); ``lib`` resolves to oracle/_ref/lib while it loads."""
    global _MODS
    if _MODS is None:
        if not available():
            raise RuntimeError("oracle/_ref is not staged (run oracle/make_ref.py in the build container)")
        sys.path.insert(0, REF_DIR)
        try:
            for name in [m for m in sys.modules if m == "lib" or m.startswith("lib.")]:
                del sys.modules[name]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                import lib.metrics as ref_metrics      # noqa
                spec = importlib.util.spec_from_file_location("_staged_reference_reward", os.path.join(REF_DIR, "reward.py"))
                ref_reward = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(ref_reward)
        finally:
            sys.path.remove(REF_DIR)
        _MODS = (ref_reward, ref_metrics)
    return _MODS


def build_cache(pk, iouv):
    """The reference's cached per-image statistics (lib/data.py:63-83) from packed arrays, TP flags by the
    reference's own ``box_correct`` (lib/metrics.py:38-64)."""
    _, metrics = modules()
    iouv = np.asarray(iouv, dtype=np.float64)
    M, T = pk.num_images, len(iouv)

    def one(off, box, cls, conf):
        out = []
        for i in range(M):
            a, b = off[i], off[i + 1]
            la, lb = pk.l_off[i], pk.l_off[i + 1]
            if a == b:
                out.append((np.zeros((0, T), dtype=bool), np.array([]), np.array([])))
            elif la == lb:
                out.append((np.zeros((b - a, T), dtype=bool), conf[a:b], cls[a:b].astype(np.int64)))
            else:
                det6 = np.column_stack([box[a:b], conf[a:b], cls[a:b]])
                lab5 = np.column_stack([pk.l_cls[la:lb], pk.l_box[la:lb]])
                out.append((metrics.box_correct(det6, lab5, iouv), conf[a:b], cls[a:b].astype(np.int64)))
        return out

    wd = one(pk.w_off, pk.w_box, pk.w_cls, pk.w_conf)
    sd = one(pk.s_off, pk.s_box, pk.s_cls, pk.s_conf)
    lc = [pk.l_cls[pk.l_off[i]:pk.l_off[i + 1]].astype(np.int64) if pk.l_off[i + 1] > pk.l_off[i] else np.array([])
          for i in range(M)]
    return wd, sd, lc


def compute_orie_seeded(i, weak_data, strong_data, labels, num_ensemble, seed):
    """The reference's ``compute_orie`` (reward.py:16-52) for target ``i`` with numpy's global generator seeded first
    (upstream is unseeded); stdout and numpy's trapz deprecation warning are swallowed."""
    reward, _ = modules()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(seed % (2 ** 32))
        return float(reward.compute_orie(int(i), weak_data, strong_data, labels, num_ensemble))


def ensemble_of(M, i, num_ensemble, seed):
    """The ensemble ``compute_orie_seeded`` draws for (i, seed): reward.py:28-38 replayed with the same generator."""
    n = max(0, min(int(num_ensemble), M - 1))
    np.random.seed(seed % (2 ** 32))
    idx = np.arange(M - 1)
    idx[i:] += 1
    return np.random.permutation(idx)[:n]


def orie_with_members(i, weak_data, strong_data, labels, members):
    """reward.py:40-50 for an EXPLICIT ensemble (``members`` = image indices, the target excluded): the same
    concatenations and the reference's own ``ap_per_class`` (lib/metrics.py:89-124), only the random draw of
    reward.py:35-38 is replaced by the given indices.  NaN (no ground truth) is returned as is."""
    _, metrics = modules()
    ens = [int(e) for e in members]
    n = len(ens)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gt = np.concatenate([labels[idx] for idx in ens] + [labels[i]]).astype(int)
        weak_img = [np.concatenate(x, axis=0) for x in zip(*([weak_data[idx] for idx in ens] + [weak_data[i]]))]
        weak_map = metrics.ap_per_class(*weak_img, gt)
        strong_img = [np.concatenate(x, axis=0) for x in zip(*([weak_data[idx] for idx in ens] + [strong_data[i]]))]
        strong_map = metrics.ap_per_class(*strong_img, gt)
        return float((np.mean(strong_map) - np.mean(weak_map)) * (n + 1))


def dcsb_all(weak_data, strong_data):
    """reward.py:55-69 for every image."""
    reward, _ = modules()
    with contextlib.redirect_stdout(io.StringIO()):
        return np.array([reward.compute_dcsb(i, weak_data, strong_data) for i in range(len(weak_data))], dtype=np.int64)
