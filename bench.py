#!/usr/bin/env python
"""ORIE rewards/sec on synthetic COCO-shaped detections (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = the whole hot path over the workload, from the packed dataset
already resident in HBM to the reward vector: FP64 IoU + TP matching for both
detectors, dataset index build (sort by class/confidence), device-side
ensemble draw, membership walk, 101-point AP integration, (N+1)*dmAP, and for
N>1 GPUs one NCCL collective (all-reduce of zero-padded per-target AP sums: the
ranks form class groups x target blocks).  Where the job can be recorded it is
replayed as ONE CUDA graph per step.  ``value`` is rewards (= target images) per
second over the K timed steps (CUDA events per step, summed; L2 flushed and, at
N>1, the ranks aligned by a barrier before each step, neither timed; max over
ranks).  ``e2e`` is the same job through the
public Python API starting from PINNED HOST buffers (H2D of the packed dataset
and D2H of the rewards inside the timed region, wall clock).

Every line carries a parity check of what was just timed (all workloads, all
N): rank 0 compares a sample of the gathered / reduced reward vector of the last
timed step with the CPU reference on the same ensembles, and the TP flags of a
sample of images bit for bit.

``--impl reference`` times the reference's own ``compute_orie`` (staged under
oracle/_ref by oracle/make_ref.py; the numpy port oracle/orie_oracle.py if the
staging is absent) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth config, method, ensemble size, IoU thresholds)           BASELINE.json configs[k]
    "coco5000": ("coco5000", "orie", 1000, 10),        # configs[1] - the metric's configuration
    "smoke500": ("smoke500", "orie", 100, 1),          # configs[0]
    "voc4952": ("voc4952", "orie", 1000, 10),          # configs[2]
    "coco5000_ori": ("coco5000", "orie", 0, 10),       # configs[3]: ORI = orie with N = 0 (reward.py:24)
    "coco5000_dcsb": ("coco5000", "dcsb", 0, 10),      # configs[3]: DCSB (reward.py:55-69) after set_data's matching
    "sweep50k": ("sweep50k", "orie", 5000, 10),        # configs[4]
}
METRIC = "ORIE rewards/sec (COCO-shape, 1000-ens)"
UNIT = "rewards/s"


def metric_name(workload):
    method, N = WORKLOADS[workload][1], WORKLOADS[workload][2]
    if workload == "coco5000":
        return METRIC
    if method == "dcsb":
        return f"DCSB rewards/sec ({workload})"
    return f"{'ORI' if N == 0 else 'ORIE'} rewards/sec ({workload}, {N}-ens)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="coco5000", choices=sorted(WORKLOADS))
    ap.add_argument("--num-images", type=int, default=0, help="override the workload's image count (debug)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="targets in the CPU baseline / parity samples (0 = by workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the timed CPU baseline (the parity check still runs)")
    ap.add_argument("--seg-chunks", type=int, default=0, help="segment length override (0 = engine default)")
    ap.add_argument("--shard", default="auto",
                    help="multi-GPU decomposition: auto | classes | targets | grid:AxB (A class groups x B target blocks); "
                         "always one all-reduce of 3 doubles per target")
    ap.add_argument("--workspace-gb", type=int, default=16, help="reward-pass workspace budget; targets run in waves that fit")
    ap.add_argument("--no-graph", action="store_true", help="time the plain calls instead of replaying a recorded job")
    return ap.parse_args()


def dataset(workload, num_images=0):
    import orie_b200  # noqa: F401
    from orie_b200 import data, synth
    cfg, method, N, T = WORKLOADS[workload]
    ds = synth.make(cfg, num_images=num_images or None)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    iouv = np.array([0.5]) if T == 1 else np.linspace(0.5, 0.95, 10)
    return ds, pk, method, N, iouv


def algorithmic_bytes(pk, N):
    """bytes(i) = 4N + sum_{e in E_i}(12 D_w(e) + 2 G(e)) + 12 (D_w(i) + D_s(i)) + 2 G(i) + 8,
    with the ensemble term taken in expectation over the random draw — the records the REFERENCE's formulation
    gathers per target."""
    M = pk.num_images
    N = max(0, min(N, M - 1))
    dw, dsn, g = np.diff(pk.w_off), np.diff(pk.s_off), np.diff(pk.l_off)
    v = 12.0 * dw + 2.0 * g
    ens = (v.sum() - v) * (N / max(M - 1, 1))
    per = 4.0 * N + ens + 12.0 * (dw + dsn) + 2.0 * g + 8.0
    return per     # float64[M]


def matching_bytes(pk):
    """sum_img 42 (D_w + D_s) + 72 G (boxes + classes in, TP mask + match index out)."""
    return 42.0 * (len(pk.w_cls) + len(pk.s_cls)) + 72.0 * len(pk.l_cls)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.perf_counter()

    def stop(self, windows=()):
        """``windows``: (t0, t1) perf_counter intervals of the timed regions; the median SM clock is taken over the
        samples that fall inside them (all samples if none does)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, sm_in, mx, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            try:
                clk = float(r[1]); mx.append(float(r[2]))
            except Exception:  # noqa: BLE001
                continue
            sm.append(clk)
            inside = any(a <= ts <= b for a, b in windows)
            if inside:
                sm_in.append(clk)
            if inside or not windows:
                for n, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
        use = sm_in or sm
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_in_timed_regions": len(sm_in), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU reference (checker / baseline arm)
def cpu_impl():
    """("reference", module) when the reference's own sources are staged under oracle/_ref, else ("port", None)."""
    from oracle import ref_local as RL
    return ("reference", RL) if RL.available() else ("port", None)


def cpu_cache_from_flags(pk, iouv, tp_w, tp_s):
    """Cached per-image statistics in the reference's layout (lib/data.py:63-83) from already-verified TP flags."""
    M, T = pk.num_images, len(iouv)

    def one(off, cls, conf, tp):
        out = []
        for i in range(M):
            a, b = off[i], off[i + 1]
            if a == b:
                out.append((np.zeros((0, T), dtype=bool), np.array([]), np.array([])))
            else:
                out.append((tp[a:b], conf[a:b], cls[a:b].astype(np.int64)))
        return out

    wd = one(pk.w_off, pk.w_cls, pk.w_conf, tp_w)
    sd = one(pk.s_off, pk.s_cls, pk.s_conf, tp_s)
    lc = [pk.l_cls[pk.l_off[i]:pk.l_off[i + 1]].astype(np.int64) if pk.l_off[i + 1] > pk.l_off[i] else np.array([])
          for i in range(M)]
    return wd, sd, lc


def cpu_flags(pk, iouv, images, strong):
    """TP flags of the given images by the CPU matcher (the reference's box_correct when staged, else the port)."""
    kind, RL = cpu_impl()
    from oracle import orie_oracle as O
    off, box, cls, conf = (pk.s_off, pk.s_box, pk.s_cls, pk.s_conf) if strong else (pk.w_off, pk.w_box, pk.w_cls, pk.w_conf)
    T = len(iouv)
    out = {}
    metrics = RL.modules()[1] if RL is not None else None
    for i in images:
        a, b = off[i], off[i + 1]
        la, lb = pk.l_off[i], pk.l_off[i + 1]
        if a == b:
            out[i] = np.zeros((0, T), dtype=bool)
        elif la == lb:
            out[i] = np.zeros((b - a, T), dtype=bool)
        elif metrics is not None:
            out[i] = metrics.box_correct(np.column_stack([box[a:b], conf[a:b], cls[a:b]]),
                                         np.column_stack([pk.l_cls[la:lb], pk.l_box[la:lb]]), iouv)
        else:
            out[i] = O.match_detections_sortunique(box[a:b], cls[a:b], pk.l_box[la:lb], pk.l_cls[la:lb], iouv)[0]
    return out


def cpu_orie_members(i, wd, sd, lc, members):
    kind, RL = cpu_impl()
    if RL is not None:
        r = RL.orie_with_members(int(i), wd, sd, lc, members)
    else:
        from oracle import orie_oracle as O
        r = O.orie_one(int(i), wd, sd, lc, members)[0]
    return 0.0 if np.isnan(r) else r


_CPU = {}


def _cpu_cache_chunk(rng):
    a, b = rng
    pk, iouv = _CPU["pk"], _CPU["iouv"]
    imgs = range(a, b)
    return a, b, cpu_flags(pk, iouv, imgs, False), cpu_flags(pk, iouv, imgs, True)


def _cpu_one(args):
    i, seed = args
    kind, RL = cpu_impl()
    if RL is not None:       # the reference's compute_orie, numpy's global generator seeded per target
        r = RL.compute_orie_seeded(i, _CPU["wd"], _CPU["sd"], _CPU["lc"], _CPU["N"], seed)
    else:
        from oracle import orie_oracle as O
        r = O.orie_one(i, _CPU["wd"], _CPU["sd"], _CPU["lc"], O.ensemble_indices(len(_CPU["lc"]), i, _CPU["N"], seed))[0]
    return r


def run_reference(args):
    """The reference's CPU implementation of the path on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    ds, pk, method, N, iouv = dataset(args.workload, args.num_images)
    kind, RL = cpu_impl()
    if RL is not None:
        RL.modules()                         # import once in the parent (torch / torchvision come with lib/data.py), workers fork
    M = pk.num_images
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    ctx = mp.get_context("fork")
    _CPU["pk"], _CPU["iouv"], _CPU["N"] = pk, iouv, N
    t0 = time.perf_counter()
    T = len(iouv)
    tp_w = np.zeros((len(pk.w_cls), T), dtype=bool)
    tp_s = np.zeros((len(pk.s_cls), T), dtype=bool)
    step_img = max(1, M // (4 * workers))
    with ctx.Pool(workers) as pool:       # set_data's matching (lib/data.py:63-83): untimed, like the reference's own timer
        for a, b, fw, fs in pool.imap_unordered(_cpu_cache_chunk, [(a, min(a + step_img, M)) for a in range(0, M, step_img)]):
            for i in range(a, b):
                tp_w[pk.w_off[i]:pk.w_off[i + 1]] = fw[i]
                tp_s[pk.s_off[i]:pk.s_off[i + 1]] = fs[i]
    _CPU["wd"], _CPU["sd"], _CPU["lc"] = cpu_cache_from_flags(pk, iouv, tp_w, tp_s)
    t_cache = time.perf_counter() - t0
    if method == "dcsb":
        sample = M
    else:
        sample = max(2 * workers, 16) if N > 0 else min(M, 2000)
    rng = np.random.default_rng(0)
    times = []
    with ctx.Pool(workers) as pool:
        for step in range(args.warmup + args.steps):
            targets = rng.choice(M, size=min(sample, M), replace=False)
            t = time.perf_counter()
            if method == "dcsb":
                (RL.dcsb_all(_CPU["wd"], _CPU["sd"]) if RL is not None else
                 __import__("oracle.orie_oracle", fromlist=["x"]).dcsb_all(_CPU["wd"], _CPU["sd"]))
            else:
                jobs = [(int(i), 10_000 * step + int(i)) for i in targets]
                pool.map(_cpu_one, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = len(targets) * len(times) / total
    what = ("compute_dcsb (reward.py:55-69), one thread" if method == "dcsb" else
            "compute_orie (reward.py:16-52) incl. its own random draw, one call per target in a process pool")
    src = ("the reference's own sources staged under oracle/_ref (oracle/make_ref.py)" if kind == "reference"
           else "oracle/orie_oracle.py (numpy port; oracle/_ref not staged)")
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "images": M, "classes": pk.num_classes, "num_ensemble": min(N, M - 1),
                   "iou_thresholds": len(iouv), "method": method,
                   "step": f"{len(targets)} sampled targets per step (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1 if method == "dcsb" else workers, "kind": kind,
                         "sample": f"{len(targets)} targets/step x {len(times)} steps on a host with {cores} cpus; {what}; {src}; "
                                   f"TP cache (set_data's matching) built untimed in {t_cache:.1f}s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ roofline
def kernel_counters(workload):
    """Per-launch ncu counters of this workload's kernels on one GPU (profiles/r02_kernel_counters.json, extracted from
    the committed ncu capture of this same command by profiles/ncu_counters.py): warp instructions executed, DRAM
    bytes, active lanes per instruction."""
    path = os.path.join(ROOT, "profiles", "r02_kernel_counters.json")
    if not os.path.exists(path):
        return {}
    try:
        return json.load(open(path)).get(workload, {})
    except Exception:  # noqa: BLE001
        return {}


def roofline_block(workload, world, dom, means, clk_mhz, info, pk, N, nt, hbm_peak, hbm_src, alg_bytes):
    """The bound the ncu counters say applies to the dominant kernel: warp-instruction ISSUE (the kernels are
    integer / bit work served from L2, DRAM traffic is a few percent of peak).  achieved = warp instructions of one
    launch (ncu smsp__inst_executed.sum of the committed capture of this workload) / the kernel's event-timed duration
    in THIS run; peak = SMs x 4 schedulers x SM clock.  Without a capture for the workload (or at N > 1, where the
    per-rank instruction count differs) the block falls back to an HBM bound on the bytes the ENGINE's formulation
    moves per launch (slot stream per 32-target batch + event records), which is always <= 1."""
    kname = {"walk_ms": "walk_kernel", "ap_ms": "ap_kernel", "match_ms": "match_kernel"}[dom]
    ms = means.get(dom, 0.0)
    sms = 148
    ctr = kernel_counters(workload).get(kname) if world == 1 else None
    nb = (nt + 31) // 32
    # the slot stream is read once per batch — per PAIR of batches where two membership tables fit shared memory (or live
    # in global memory) — as one packed 32-bit word per slot up to 65535 images, else image u32 + TP mask u16
    table = ((pk.num_images + 1 + 31) // 32) * 32 * 4
    pairs = table > 112 * 1024 or 2 * table <= 56 * 1024
    slot_stream = float((nb + 1) // 2 if pairs else nb) * float(info.get("slots", 0)) * (4.0 if pk.num_images <= 65535 else 6.0)
    ev_touch = float(nt) * float(info.get("events", 0)) * 4.0 * (min(N, pk.num_images - 1) + 1) / max(pk.num_images, 1)
    seg_tab = float(info.get("segments", 0)) * float(nb * 32) * 8.0
    engine_bytes = {"walk_ms": slot_stream + ev_touch + seg_tab, "ap_ms": ev_touch + seg_tab,
                    "match_ms": matching_bytes(pk)}[dom]
    hbm_engine = {"bytes_per_launch": engine_bytes, "achieved": engine_bytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0,
                  "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src}
    hbm_engine["frac"] = hbm_engine["achieved"] / hbm_peak
    naive = {"bytes_per_launch": alg_bytes, "achieved": alg_bytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0, "peak": hbm_peak,
             "unit": "GB/s", "note": "SURVEY 8d bytes: the records the REFERENCE's formulation gathers per target; the engine "
                                     "never moves them, so this ratio can exceed 1 and is not a roofline"}
    naive["ratio"] = naive["achieved"] / hbm_peak
    if ctr and ms > 0 and clk_mhz:
        peak = sms * 4 * clk_mhz * 1e6 / 1e9                      # G warp-inst/s
        ach = ctr["inst"] / (ms / 1e3) / 1e9
        out = {"bound": "issue", "kernel": kname, "achieved": ach, "peak": peak, "unit": "Gwarp-inst/s", "frac": ach / peak,
               "traffic": ctr.get("dram_bytes"), "lanes_per_inst": ctr.get("lanes_per_inst"),
               "inst_per_launch": ctr["inst"], "inst_source": "profiles/r02_kernel_counters.json (ncu smsp__inst_executed.sum)",
               "peak_source": f"{sms} SMs x 4 schedulers x {clk_mhz:.0f} MHz (median SM clock sampled in the timed region)"}
    else:
        out = {"bound": "hbm", "kernel": kname, "achieved": hbm_engine["achieved"], "peak": hbm_peak, "unit": "GB/s",
               "frac": hbm_engine["frac"], "traffic": (ctr or {}).get("dram_bytes"), "peak_source": hbm_src,
               "note": "no ncu instruction count committed for this workload / rank count: bytes = what the engine's "
                       "formulation moves per launch (slot stream per 32-target batch, event records, segment tables)"}
    out["kernel_ms"] = means
    out["hbm_engine_bytes"] = hbm_engine
    out["vs_naive_streaming"] = naive
    return out


# ------------------------------------------------------------------ B200 arm
def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import orie_b200  # noqa: F401
    from orie_b200 import _lib
    from orie_b200.engine import (DevicePacked, Engine, HostPacked, ReplayJob, class_shard, combine_sums, rewards_from_sums, shard_of_rank,
                                  shard_plan)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    ds, pk, method, N, iouv = dataset(args.workload, args.num_images)
    M, T = pk.num_images, len(iouv)
    Nc = max(0, min(N, M - 1))
    is_dcsb = method == "dcsb"
    # class groups x target blocks (engine.shard_plan); one all-reduce of zero-padded per-target sums combines the ranks
    sharded = world > 1 and not is_dcsb
    rc_n, rt_n = shard_plan(M, world, args.shard, pk.num_classes) if sharded else (1, 1)
    rc, _, t0, nt = shard_of_rank(rank, M, world, args.shard, pk.num_classes) if sharded else (0, 1, 0, M)
    pk_all = pk
    shard_s = 0.0
    if rc_n > 1:
        ts = time.perf_counter()
        pk = class_shard(pk_all, rc, rc_n)           # this rank's classes, all images (host-side partition at pack time)
        shard_s = time.perf_counter() - ts
    hp = HostPacked(pk)                      # pinned once, outside every timed region
    dp = DevicePacked(hp, dev)               # resident in HBM for the `value` steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    gathered_i = torch.zeros(M, dtype=torch.int64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms = {"label_walk_ms": [], "walk_ms": [], "ap_ms": [], "finalize_ms": []}
    phase_ms = {"match_index_ms": [], "reward_ms": []}

    def step(seed, record, profile=False):
        """match + index + ensemble draw + rewards (+ collective) from HBM-resident inputs."""
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng = Engine(dp, iouv=iouv, seg_chunks=args.seg_chunks, index=not is_dcsb)
        e1.record()
        s = C.c_void_p(eng.stream.cuda_stream)
        if is_dcsb:
            _lib.check(lib.orie_dcsb(C.c_void_p(eng.w_conf.data_ptr()), C.c_void_p(eng.w_off.data_ptr()),
                                     C.c_void_p(eng.s_conf.data_ptr()), C.c_void_p(eng.s_off.data_ptr()), M,
                                     C.c_void_p(gathered_i.data_ptr()), s))
            e2.record()
            e2.synchronize()
            if record:
                phase_ms["match_index_ms"].append(e0.elapsed_time(e1))
                phase_ms["reward_ms"].append(e1.elapsed_time(e2))
            eng.close()
            return e0.elapsed_time(e2), {}
        wave, ws_bytes = eng.plan_waves(nt, args.workspace_gb << 30)      # no synchronisation when the bound fits the budget
        ws = eng._workspace(ws_bytes)
        bits = torch.empty((max(wave, 1), eng.ens_words), dtype=torch.int32, device=dev)
        mine = torch.zeros(max(nt, 1), dtype=torch.float64, device=dev)
        sums = torch.zeros((max(nt, 1), 3), dtype=torch.float64, device=dev) if sharded else None
        ms = [0.0] * 4
        for a in range(0, nt, max(wave, 1)):            # target waves sized by the workspace budget
            cnt = min(wave, nt - a)
            part = (C.c_float * 4)()
            _lib.check(lib.orie_ensemble_sample(eng._handle, t0 + a, cnt, Nc, seed, C.c_void_p(bits.data_ptr()), s))
            out_r = C.c_void_p(0) if sharded else C.c_void_p(mine.data_ptr() + 8 * a)
            out_s = C.c_void_p(sums.data_ptr() + 24 * a) if sharded else C.c_void_p(0)
            if profile:      # untimed passes only: events around every kernel, synchronises
                _lib.check(lib.orie_reward_profile(eng._handle, t0 + a, cnt, C.c_void_p(bits.data_ptr()), Nc,
                                                   C.c_void_p(ws.data_ptr()), ws.numel(), out_r, out_s, 0, s, part))
                ms = [x + float(y) for x, y in zip(ms, part)]
            elif sharded:
                _lib.check(lib.orie_reward_sums(eng._handle, t0 + a, cnt, C.c_void_p(bits.data_ptr()), Nc,
                                                C.c_void_p(ws.data_ptr()), ws.numel(), out_s, 0, s))
            else:
                _lib.check(lib.orie_reward(eng._handle, t0 + a, cnt, C.c_void_p(bits.data_ptr()), Nc,
                                           C.c_void_p(ws.data_ptr()), ws.numel(), out_r, C.c_void_p(0), s))
        # 3 doubles per target: AP sums are additive over classes
        result["rewards"] = combine_sums(sums[:nt], t0, M, T, Nc) if sharded else mine[:M]
        e2.record()
        e2.synchronize()
        if profile:
            for k, v in zip(("label_walk_ms", "walk_ms", "ap_ms", "finalize_ms"), ms):
                kernel_ms[k].append(float(v))
            if job is not None:              # the timed steps were graph replays: phase split from these plain-call passes
                phase_ms["match_index_ms"].append(e0.elapsed_time(e1))
        elif record:
            phase_ms["match_index_ms"].append(e0.elapsed_time(e1))
            phase_ms["reward_ms"].append(e1.elapsed_time(e2))
        info = eng.info                      # after the timed region: the exact sizes (the build itself never synchronised)
        eng.check_status()
        eng.close()
        return e0.elapsed_time(e2), info

    # The timed steps replay ONE recorded job (CUDA graph: matching + index build + draw + walk + AP, this rank's
    # share) per step, followed by the collective.  Datasets whose reward pass has to run in waves sized from
    # device-side facts (the 50k sweep) cannot be recorded and go through the plain calls.
    job = None
    if not is_dcsb and not args.no_graph:
        try:
            job = ReplayJob(dp, iouv=iouv, num_ensemble=N, t0=t0, nt=nt, sums=sharded, total_images=M,
                            tuning=dict(seg_chunks=args.seg_chunks), workspace_budget=args.workspace_gb << 30)
        except RuntimeError as e:
            if "cannot be recorded" not in str(e):
                raise
    replay_info = {}
    result = {}

    def replay_step(seed, record):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = job.run(seed)
        result["rewards"] = combine_sums(out, t0, M, T, Nc) if sharded else out       # device tensor f64[M]
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1), replay_info

    timed_step = replay_step if job is not None else step
    clocks = ClockSampler(local)
    windows = []
    if rank == 0:
        clocks.start()                      # sampled from the warm-up to the end of the e2e runs
    for w in range(args.warmup):
        flush.fill_(w)
        timed_step(1000 + w, False)
    # small workloads: the timed region can be shorter than the sampler's start-up — keep warming up (untimed) until it
    # has delivered its first sample, so that the clock record is taken under this job's load
    t_wait = time.perf_counter()
    while rank == 0 and clocks.proc is not None and not clocks.rows and time.perf_counter() - t_wait < 3.0:
        timed_step(1500, False)
    barrier()
    launches0 = lib.orie_launch_count()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    info = None
    for k in range(args.steps):
        flush.fill_(k)                      # L2 flush, not timed
        barrier()                           # the ranks start a step together (not timed): without it every rank's interval
        #                                     also holds its wait, inside the all-reduce, for the rank that started last
        ms, info = timed_step(2000 + k, True)
        dev_ms += ms
    last_seed = 2000 + args.steps - 1
    barrier()
    wall = time.perf_counter() - wall0
    windows.append((wall0, wall0 + wall))
    launches = lib.orie_launch_count() - launches0
    if job is not None:
        launches += job.launches_per_replay * args.steps        # kernels inside the replayed graph (counted when it was recorded)
        job.check_status()
        info = dict(job.engine.info)
    result_last = (gathered_i if is_dcsb else result["rewards"]).cpu().numpy().copy()       # what the last timed step produced
    if not is_dcsb:
        for k in range(min(args.steps, 5)):     # per-kernel durations for the roofline: same step, events around each kernel, untimed
            flush.fill_(k)
            torch.cuda.synchronize()
            step(2000 + k, False, profile=True)
    # matching alone (both detectors' orie_match launches on one stream), event-timed, for the matcher's roofline
    match_ms = []
    eng = Engine(dp, iouv=iouv, index=False)
    torch.cuda.synchronize()
    for k in range(5):
        flush.fill_(k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng._match(torch.cuda.current_stream())
        b.record()
        b.synchronize()
        match_ms.append(a.elapsed_time(b))
    eng.close()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = M * args.steps / (dev_ms / 1e3)

    # ---- e2e: pinned host buffers -> rewards on the host, through the public API
    e2e_times = []
    e2e_warm = max(3, args.warmup)           # the first passes pay for pinned staging / allocator pools, like any warm-up
    e2e_t0 = time.perf_counter()
    for k in range(e2e_warm + max(2, min(args.steps, 10))):
        barrier()
        t_start = time.perf_counter()
        eng = Engine(hp, iouv=iouv, device=dev, index=not is_dcsb)
        if is_dcsb:
            host = eng.dcsb()
        elif sharded:
            sums = eng.orie_sums_device(N, seed=3000 + k, t0=t0, nt=nt, total_images=M)
            eng.stream.synchronize()
            host = combine_sums(sums, t0, M, T, Nc).cpu()
        else:
            host = eng.orie_device(N, seed=3000 + k).cpu()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t_start
        eng.close()
        if k >= e2e_warm:
            e2e_times.append(dt)
    windows.append((e2e_t0, time.perf_counter()))
    t = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = M * len(e2e_times) / float(t.item())
    clk = clocks.stop(windows) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.barrier()                   # rank 0 is still checking parity
            dist.destroy_process_group()
        return

    # ---- parity of what was just timed (every workload, every N), against the CPU reference on rank 0
    kind, RL = cpu_impl()
    S = args.cpu_sample or {"sweep50k": 8}.get(args.workload, 32)
    S = min(S, M)
    parity = {"checker": "oracle/_ref (the reference's own box_correct / ap_per_class / compute_orie)" if kind == "reference"
              else "oracle/orie_oracle.py (numpy port)"}
    eng = Engine(DevicePacked(pk_all, dev), iouv=iouv, index=not is_dcsb)     # un-sharded, for TP flags and ensemble bitmaps
    wtp, stp, _, _ = eng.tp_flags()
    imgs = sorted(set(np.linspace(0, M - 1, min(M, 200)).astype(np.int64).tolist()))
    bad = 0
    for strong, tp, off in ((False, wtp, pk_all.w_off), (True, stp, pk_all.s_off)):
        want = cpu_flags(pk_all, iouv, imgs, strong)
        for i in imgs:
            bad += int(np.count_nonzero(tp[off[i]:off[i + 1]] != want[i]))
    parity["tp_flag_images_checked"] = 2 * len(imgs)
    parity["tp_flag_mismatches"] = bad
    wd, sd, lc = cpu_cache_from_flags(pk_all, iouv, wtp, stp)
    cpu = None
    if is_dcsb:
        want = RL.dcsb_all(wd, sd) if RL is not None else __import__("oracle.orie_oracle", fromlist=["x"]).dcsb_all(wd, sd)
        parity["targets_checked"] = M
        parity["max_abs_err"] = float(np.abs(result_last - want).max())
        tc = time.perf_counter()
        (RL.dcsb_all(wd, sd) if RL is not None else None)
        tc = time.perf_counter() - tc
        if RL is not None and not args.no_cpu_baseline and world == 1:
            cpu = {"value": M / tc, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"compute_dcsb (reward.py:55-69) for all {M} images, one thread, matching excluded"}
    else:
        # (1) the distributed result of the last timed step, device-drawn ensembles replayed on the CPU
        targets = np.unique(np.linspace(0, M - 1, S).astype(np.int64))
        errs = []
        for i in targets:
            b0 = int(i) // 32 * 32
            bits = eng.sample_bits(N, seed=last_seed, t0=b0, nt=min(32, M - b0))[int(i) - b0]
            member = np.nonzero(((bits[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).reshape(-1)[:M])[0]
            errs.append(abs(cpu_orie_members(int(i), wd, sd, lc, member) - result_last[int(i)]))
        parity["targets_checked"] = int(len(targets))
        parity["max_abs_err"] = float(max(errs))
        parity["what"] = (f"reward vector of the last timed step (after the collective, {world} gpu(s)) at {len(targets)} evenly "
                          "spaced targets vs the CPU reference on the same device-drawn ensembles")
        # (2) N = 1 only: the reference's compute_orie itself, timed (cpu_baseline) and compared on its own numpy ensembles
        if world == 1 and not args.no_cpu_baseline:
            b0 = (M // 64) * 32
            nb = min(S, 32, M - b0)
            seeds = [777_000 + b0 + r for r in range(nb)]
            if RL is not None:
                em = np.stack([RL.ensemble_of(M, b0 + r, N, seeds[r]) for r in range(nb)]).astype(np.int32)
            else:
                from oracle import orie_oracle as O
                em = np.stack([O.ensemble_indices(M, b0 + r, N, seeds[r]) for r in range(nb)]).astype(np.int32)
            got = eng.orie(N, ens_matrix=em, t0=b0, nt=nb)
            _CPU.update(wd=wd, sd=sd, lc=lc, N=N)
            tc = time.perf_counter()
            want = np.array([_cpu_one((b0 + r, seeds[r])) for r in range(nb)])
            tc = time.perf_counter() - tc
            want = np.where(np.isnan(want), 0, want)
            parity["compute_orie_targets_checked"] = nb
            parity["compute_orie_max_abs_err"] = float(np.abs(got - want).max())
            cpu = {"value": nb / tc, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"{nb} consecutive targets of the same workload, one thread, "
                             + ("the reference's compute_orie (reward.py:16-52, staged in oracle/_ref) seeded per target"
                                if kind == "reference" else "oracle/orie_oracle.py")
                             + f" (reward phase only, TP cache prebuilt) on a host with {os.cpu_count()} cpus"}
    eng.close()
    parity_err = parity.get("max_abs_err")

    # ---- roofline of the dominant kernel (CUDA events around each kernel, averaged over the profiled steps)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    means = {k: (sum(v) / len(v) if v else 0.0) for k, v in kernel_ms.items()}
    means["match_ms"] = statistics.median(match_ms)      # the median drops a pass that paid for a fresh allocator block
    dom = "match_ms" if is_dcsb else max(("walk_ms", "ap_ms", "match_ms"), key=lambda k: means[k])
    per_target = algorithmic_bytes(pk, Nc)      # pk = this rank's share of the records when classes are sharded
    alg_bytes = float(per_target[t0:t0 + nt].sum()) if dom != "match_ms" else matching_bytes(pk)
    roofline = roofline_block(args.workload, world, dom, means, (clk or {}).get("sm_mhz"), info or {}, pk, Nc, nt, peak, peak_src,
                              alg_bytes)

    line = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "method": method, "images": M, "classes": pk_all.num_classes, "num_ensemble": Nc,
                   "iou_thresholds": T, "weak_dets": int(len(pk_all.w_cls)), "strong_dets": int(len(pk_all.s_cls)),
                   "labels": int(len(pk_all.l_cls)),
                   "parallelism": ("one gpu" if world == 1 else
                                   f"{rc_n} class group(s) x {rt_n} target block(s) over {world} gpus (a rank: the rows of its "
                                   f"classes for all images, the targets of its block), one all-reduce of 3 doubles per target"),
                   "step": ("TP matching (2 detectors) + DCSB count" if is_dcsb else
                            "TP matching (2 detectors) + index build + ensemble draw + membership walk + AP + collective"),
                   "launch": ("one CUDA-graph replay per step (engine.ReplayJob), then the collective" if job is not None
                              else "plain C-ABI calls"),
                   "l2": "flushed between steps (256 MiB write, not timed)", "ensembles": "device-side Philox draw, seed per step",
                   "index": {k: info[k] for k in ("slots", "segments", "events", "seg_chunks", "class_groups")} if info else None,
                   "phase_ms": {k: sum(v) / len(v) for k, v in phase_ms.items() if v},
                   "phase_ms_from": "plain-call passes after the timed region" if job is not None else "the timed steps",
                   "wall_s_timed_region": wall,
                   "host_class_shard_s": shard_s if rc_n > 1 else None},
        "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pk.nbytes()), "d2h_bytes_per_step": int(M * 8),
                "note": "bytes per rank" if world > 1 else "",
                "ms_per_step": 1e3 * sum(e2e_times) / len(e2e_times), "timer": "wall clock around Engine(pinned host) + rewards + .cpu()"},
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "parity_max_abs_err_vs_oracle": parity_err,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
