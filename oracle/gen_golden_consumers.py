"""Freeze golden vectors for the two consumers of the reward vector from the LIVE reference:

    python oracle/gen_golden_consumers.py      (build container only; writes tests/golden/dcsb_fit*.npz, rank_norm.npz)

* ``baseline.fit_dcsb`` (baseline.py:67-152) is imported from /root/reference and called with the inputs
  baseline.py:161-206 builds for ``--baseline dcsb``; its thresholds are read back from the pickle it saves.
* the rank normalisation is not a function upstream: regression.py:439-441 are three lines inside ``main``.  They are
  read from the reference's source file and executed verbatim on the fixture's train / validation rewards.
"""
import contextlib
import importlib.util
import io
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import orie_b200  # noqa: E402,F401
from orie_b200 import data, synth  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def upstream_baseline():
    with R._ref_path():
        spec = importlib.util.spec_from_file_location("_upstream_baseline", os.path.join(R.REF_ROOT, "baseline.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


def dcsb_case(name, config, M, seed, folds, reward_seed):
    base = upstream_baseline()
    ds = synth.make(config, num_images=M, seed=seed, empty_det_frac=0.04)
    rng = np.random.default_rng(reward_seed)
    # upstream searches area thresholds in [0.2, 0.9): give the weak boxes areas that spread over that range, and keep
    # only a handful of confident rows per image, so that both thresholds of the grid matter
    ds.weak.rows[:, 3:5] = np.clip(ds.weak.rows[:, 3:5] * rng.uniform(2.0, 6.0, size=(len(ds.weak.rows), 1)), 0.05, 0.97)
    ds.weak.rows[:, 5] = np.where(rng.random(len(ds.weak.rows)) < 0.8, ds.weak.rows[:, 5] * 0.3, ds.weak.rows[:, 5])
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    # rewards that follow a DCSB-like rule (so that the grid search has a non-trivial optimum), with 12 % label noise
    n_t, a_t = int(rng.integers(2, 7)), float(rng.uniform(0.3, 0.7))
    reward = np.empty(M)
    for i in range(M):
        a, b = pk.w_off[i], pk.w_off[i + 1]
        conf = pk.w_conf[a:b]
        area = (pk.w_box[a:b, 2] - pk.w_box[a:b, 0]) * (pk.w_box[a:b, 3] - pk.w_box[a:b, 1])
        sel = conf > 0.25
        num, amin = int(sel.sum()), (float(area[sel].min()) if sel.any() else 0.0)
        hard = num != int((conf > 0.5).sum()) and (num > n_t or amin < a_t)
        reward[i] = (1.0 if hard else -1.0) * rng.uniform(0.05, 2.0)
    flip = rng.random(M) < 0.12
    reward[flip] *= -1
    fold = rng.permutation(M) % folds
    out = dict(l_off=ds.labels.off, l_rows=ds.labels.rows, w_off=ds.weak.off, w_rows=ds.weak.rows,
               s_off=ds.strong.off, s_rows=ds.strong.rows, reward=reward, fold=fold.astype(np.int64))
    with tempfile.TemporaryDirectory() as d:
        w, s, l = synth.write_dirs(ds, d)
        _, _, rdata = R.modules()
        names = sorted(os.listdir(l))
        names = ['.'.join(n.split('.')[:-1]) for n in names]
        weak_data = rdata.load_data(w, names, True)                                  # baseline.py:176-181
        feature = [(np.array([]), np.array([])) if len(wd) == 0 else (wd[2], base.get_area(wd[1])) for wd in weak_data]
        labels = rdata.load_data(l, names)
        label_num = np.array([0 if len(x) == 0 else len(x[0]) for x in labels], dtype=int)
        r01 = np.where(reward > 0, 1, 0)                                             # baseline.py:163-166
        for k in range(folds):
            val = fold == k
            opts = base._SaveOPT
            opts.save, opts.load, opts.model_dir, opts.model_idx = True, False, os.path.join(d, "model"), k + 1
            with contextlib.redirect_stdout(io.StringIO()):
                res = base.fit_dcsb(([f for f, v in zip(feature, val) if not v], [f for f, v in zip(feature, val) if v],
                                     r01[~val], r01[val]), label_num[~val], opts)
            conf_t, num_t, area_t = pickle.load(open(os.path.join(d, "model", f"wts{k + 1}.pickle"), "rb"))
            out[f"train_est{k}"], out[f"val_est{k}"] = res["train_est"], res["val_est"]
            out[f"model{k}"] = np.array([conf_t, num_t, area_t], dtype=np.float64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", [out[f"model{k}"].tolist() for k in range(folds)])


def rank_norm_case():
    src = open(os.path.join(R.REF_ROOT, "regression.py")).read().splitlines()
    lines = src[439:441]                    # regression.py:440-441 (1-based): the two statements under `if opts.normalize:` (:439)
    assert "val_reward = np.array([np.sum(train_reward <= x)" in lines[0] and "np.argsort(np.argsort(train_reward))" in lines[1], lines
    code = "\n".join(x.strip() for x in lines[:2])
    rng = np.random.default_rng(11)
    out = {}
    for k, (M, frac) in enumerate([(1, 0.0), (9, 0.3), (700, 0.2), (6000, 0.25)]):
        reward = rng.normal(size=M) * 10.0 ** rng.integers(-6, 3, size=M)            # tie-free
        val = rng.random(M) < frac
        if val.all():
            val[0] = False
        env = {"np": np, "train_reward": reward[~val], "val_reward": reward[val]}
        exec(code, env)                     # the reference's own two statements
        want = np.empty(M)
        want[~val], want[val] = env["train_reward"], env["val_reward"]
        out[f"reward{k}"], out[f"val{k}"], out[f"want{k}"] = reward, val, want
    out["cases"] = np.array(k + 1)
    path = os.path.join(OUT, "rank_norm.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    assert R.available(), "needs /root/reference"
    dcsb_case("dcsb_fit_coco", "smoke500", 120, 71, 3, 5)
    dcsb_case("dcsb_fit_voc", "voc4952", 36, 72, 2, 6)
    rank_norm_case()
