"""Drive the LIVE upstream implementation (read-only at /root/reference).

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the GPU box
has no /root/reference, so nothing marked ``gpu``, ``smoke()`` or ``bench.py``
may import this module.  It never copies upstream code; it imports
``reward.compute_orie`` / ``reward.compute_dcsb`` / ``lib.metrics.*`` /
``lib.data.load_data`` and replicates only the glue upstream hard-codes
(``lib/data.py:61`` fixes ``iouv=[0.5]`` as a local, so the T=10 cache is
built by calling upstream's own T-generic ``box_correct`` per image, following
``lib/data.py:63-83``).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np

REF_ROOT = os.environ.get("ORIE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "reward.py"))


@contextlib.contextmanager
def _ref_path():
    sys.path.insert(0, REF_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)


_MODS = None


def modules():
    """(reward, lib.metrics, lib.data) of the live reference.  ``reward.py`` is
    loaded by path under the name ``_upstream_reward`` because this repository
    has a ``reward.py`` of its own (the drop-in CLI)."""
    global _MODS
    if _MODS is None:
        import importlib.util
        with _ref_path():
            import lib.metrics as ref_metrics      # noqa
            import lib.data as ref_data            # noqa
            spec = importlib.util.spec_from_file_location("_upstream_reward", os.path.join(REF_ROOT, "reward.py"))
            ref_reward = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref_reward)
        _MODS = (ref_reward, ref_metrics, ref_data)
    return _MODS


def ref_set_data(weak_dir, strong_dir, label_dir, iouv=None):
    """Upstream ``set_data`` (iouv=None) or its T-generic twin."""
    _, metrics, data = modules()
    if iouv is None:
        return data.set_data(weak_dir, strong_dir, label_dir)
    iouv = np.asarray(iouv, dtype=np.float64)
    names = sorted(os.listdir(label_dir))
    names = ['.'.join(n.split('.')[:-1]) for n in names]
    weak = data.load_data(weak_dir, names, True)
    strong = data.load_data(strong_dir, names, True)
    labels = data.load_data(label_dir, names)
    T = len(iouv)

    def cached(det, lab):
        # empty-case shapes as upstream: (bool[0,T], f64[0], f64[0])
        if len(det) == 0:
            return np.zeros((0, T), dtype=bool), np.array([]), np.array([])
        cls, box, conf = det
        if len(lab) == 0:
            return np.zeros((len(cls), T), dtype=bool), conf, cls
        det6 = np.column_stack([box, conf, cls])        # x1 y1 x2 y2 conf cls
        lab5 = np.column_stack([lab[0], lab[1]])        # cls x1 y1 x2 y2
        return metrics.box_correct(det6, lab5, iouv), conf, cls

    for i in range(len(labels)):
        lab = labels[i]
        weak[i] = cached(weak[i], lab)
        strong[i] = cached(strong[i], lab)
        labels[i] = lab[0] if len(lab) > 0 else np.array([])
    return weak, strong, labels


def ref_orie(weak_data, strong_data, labels, num_ensemble, base_seed, targets=None):
    """Sequential upstream ``compute_orie`` with ``np.random.seed(base+idx)``
    before each call (upstream itself is unseeded and threaded)."""
    reward, _, _ = modules()
    targets = range(len(labels)) if targets is None else targets
    out = np.empty(len(targets), dtype=np.float64)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r, i in enumerate(targets):
            np.random.seed(base_seed + int(i))
            out[r] = reward.compute_orie(int(i), weak_data, strong_data, labels, num_ensemble)
    return out


def ref_dcsb(weak_data, strong_data):
    reward, _, _ = modules()
    with contextlib.redirect_stdout(io.StringIO()):
        return np.array([reward.compute_dcsb(i, weak_data, strong_data) for i in range(len(weak_data))], dtype=int)


def ref_box_correct(dets, labs, iouv):
    _, metrics, _ = modules()
    return metrics.box_correct(dets, labs, np.asarray(iouv, dtype=np.float64))


def ref_ap_per_class(tp, conf, pred_cls, target_cls):
    _, metrics, _ = modules()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return metrics.ap_per_class(tp, conf, pred_cls, target_cls)


def ref_test_map(weak_dir, strong_dir, label_dir, estimate_dirs, dataset_split):
    """Upstream test.py:test_map on upstream's own set_data cache (T = 1, as shipped)."""
    import importlib.util
    with _ref_path():
        spec = importlib.util.spec_from_file_location("_upstream_test", os.path.join(REF_ROOT, "test.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _, _, data = modules()
    weak_data, strong_data, labels = data.set_data(weak_dir, strong_dir, label_dir)
    labels = np.concatenate(labels).astype(int)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return mod.test_map(weak_data, strong_data, labels, list(estimate_dirs), dataset_split)
