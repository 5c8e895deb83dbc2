#!/usr/bin/env python
"""ORIE rewards/sec on synthetic COCO-shaped detections (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = the whole hot path over the workload, from the packed dataset
already resident in HBM to the reward vector: FP64 IoU + TP matching for both
detectors, dataset index build (radix sort by class/confidence), device-side
ensemble draw, membership walk, 101-point AP integration, (N+1)*dmAP, and for
N>1 GPUs one NCCL collective (all-reduce of per-target AP sums when the classes
are sharded over the ranks, all-gather of reward slices when the targets are).  ``value`` is
rewards (= target images) per second over the K timed steps (CUDA events per
step, summed; max over ranks).  ``e2e`` is the same job through the public
Python API starting from PINNED HOST buffers (H2D of the packed dataset and
D2H of the rewards inside the timed region, wall clock).

``--impl reference`` times the CPU port of the reference's algorithm
(oracle/orie_oracle.py; the reference itself is Python and does not travel to
the GPU box) on the host cores, on a bounded sample of the same workload.
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
    # name: (synth config, ensemble size, IoU thresholds)
    "coco5000": ("coco5000", 1000, 10),
    "voc4952": ("voc4952", 1000, 10),
    "smoke500": ("smoke500", 100, 1),
    "sweep50k": ("sweep50k", 5000, 10),
}
METRIC = "ORIE rewards/sec (COCO-shape, 1000-ens)"
UNIT = "rewards/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="coco5000", choices=sorted(WORKLOADS))
    ap.add_argument("--num-images", type=int, default=0, help="override the workload's image count (debug)")
    ap.add_argument("--cpu-sample", type=int, default=32, help="targets in the single-thread CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seg-chunks", type=int, default=0, help="segment length override (0 = engine default)")
    ap.add_argument("--shard", default="auto", choices=["auto", "classes", "targets"],
                    help="multi-GPU decomposition: classes (whole pipeline shrinks per rank, one all-reduce of 3 doubles "
                         "per target) or targets (index replicated, one all-gather of reward slices)")
    ap.add_argument("--workspace-gb", type=int, default=16, help="reward-pass workspace budget; targets run in waves that fit")
    return ap.parse_args()


def dataset(workload, num_images=0):
    import orie_b200  # noqa: F401
    from orie_b200 import data, synth
    cfg, N, T = WORKLOADS[workload]
    ds = synth.make(cfg, num_images=num_images or None)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    iouv = np.array([0.5]) if T == 1 else np.linspace(0.5, 0.95, 10)
    return ds, pk, N, iouv


def algorithmic_bytes(pk, N):
    """bytes(i) = 4N + sum_{e in E_i}(12 D_w(e) + 2 G(e)) + 12 (D_w(i) + D_s(i)) + 2 G(i) + 8,
    with the ensemble term taken in expectation over the random draw."""
    M = pk.num_images
    N = max(0, min(N, M - 1))
    dw, dsn, g = np.diff(pk.w_off), np.diff(pk.s_off), np.diff(pk.l_off)
    v = 12.0 * dw + 2.0 * g
    ens = (v.sum() - v) * (N / max(M - 1, 1))
    per = 4.0 * N + ens + 12.0 * (dw + dsn) + 2.0 * g + 8.0
    return per     # float64[M]


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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:  # noqa: BLE001
                continue
            for n, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU port (reference arm / baseline)
_CPU = {}


def _cpu_one(args):
    i, ens = args
    from oracle import orie_oracle as O
    return O.orie_one(i, _CPU["wd"], _CPU["sd"], _CPU["lc"], ens)[0]


def cpu_cache_from_packed(pk, iouv, tp_w=None, tp_s=None):
    """Cached per-image statistics in the reference's layout.  TP flags come from
    the CPU matcher unless already-verified flags are handed in."""
    from oracle import orie_oracle as O
    M, T = pk.num_images, len(iouv)

    def one(off, box, cls, conf, tp):
        out = []
        for i in range(M):
            a, b = off[i], off[i + 1]
            if a == b:
                out.append((np.zeros((0, T), dtype=bool), np.array([]), np.array([])))
                continue
            if tp is not None:
                flags = tp[a:b]
            else:
                la, lb = pk.l_off[i], pk.l_off[i + 1]
                flags = O.match_detections_sortunique(box[a:b], cls[a:b], pk.l_box[la:lb], pk.l_cls[la:lb], iouv)[0]
            out.append((flags, conf[a:b], cls[a:b].astype(np.int64)))
        return out

    wd = one(pk.w_off, pk.w_box, pk.w_cls, pk.w_conf, tp_w)
    sd = one(pk.s_off, pk.s_box, pk.s_cls, pk.s_conf, tp_s)
    lc = [pk.l_cls[pk.l_off[i]:pk.l_off[i + 1]].astype(np.int64) if pk.l_off[i + 1] > pk.l_off[i] else np.array([])
          for i in range(M)]
    return wd, sd, lc


def run_reference(args):
    """CPU port of the reference on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import orie_oracle as O
    ds, pk, N, iouv = dataset(args.workload, args.num_images)
    M = pk.num_images
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    t0 = time.perf_counter()
    _CPU["wd"], _CPU["sd"], _CPU["lc"] = cpu_cache_from_packed(pk, iouv)
    t_cache = time.perf_counter() - t0
    sample = max(2 * workers, 16)
    rng = np.random.default_rng(0)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(workers) as pool:
        for step in range(args.warmup + args.steps):
            targets = rng.choice(M, size=min(sample, M), replace=False)
            jobs = [(int(i), O.ensemble_indices(M, int(i), N, 10_000 * step + int(i))) for i in targets]
            t = time.perf_counter()
            pool.map(_cpu_one, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = len(targets) * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "images": M, "classes": pk.num_classes, "num_ensemble": N,
                   "iou_thresholds": len(iouv), "step": f"{len(targets)} sampled targets per step (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{len(targets)} targets/step x {len(times)} steps, process pool of {workers} on {cores} host cpus; "
                                   f"oracle/orie_oracle.py (numpy port of reward.py:16-52 + lib/metrics.py:89-148); "
                                   f"TP cache built untimed in {t_cache:.1f}s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import orie_b200  # noqa: F401
    from orie_b200 import _lib
    from orie_b200.engine import DevicePacked, Engine, HostPacked, class_shard, pick_shard, rewards_from_sums, shard_range

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

    ds, pk, N, iouv = dataset(args.workload, args.num_images)
    M, T = pk.num_images, len(iouv)
    by_class = world > 1 and pick_shard(pk.num_images, args.shard) == "classes"
    pk_all = pk
    if by_class:
        pk = class_shard(pk_all, rank, world)        # this rank's classes, all images (host-side partition, untimed)
    hp = HostPacked(pk)                      # pinned once, outside every timed region
    dp = DevicePacked(hp, dev)               # resident in HBM for the `value` steps
    t0, nt = (0, M) if by_class else shard_range(M, rank, world)
    per = M if by_class else shard_range(M, 0, world)[1]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    gathered = torch.empty(M if by_class else per * world, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms = {"label_walk_ms": [], "walk_ms": [], "ap_ms": [], "finalize_ms": []}
    phase_ms = {"match_index_ms": [], "reward_ms": []}

    def step(seed, record, profile=False):
        """match + index + ensemble draw + rewards (+ all-gather) from HBM-resident inputs."""
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng = Engine(dp, iouv=iouv, seg_chunks=args.seg_chunks)
        e1.record()
        wave = eng.wave_size(nt, args.workspace_gb << 30) if nt > 0 else 0
        ws = eng._workspace(eng.workspace_bytes(wave))
        bits = torch.empty((max(wave, 1), eng.info["ens_words"]), dtype=torch.int32, device=dev)
        mine = torch.zeros(per, dtype=torch.float64, device=dev)
        sums = torch.zeros((M, 3), dtype=torch.float64, device=dev) if by_class else None
        import ctypes as C
        s = C.c_void_p(eng.stream.cuda_stream)
        ms = [0.0] * 4
        Nc = min(N, M - 1)
        for a in range(0, nt, max(wave, 1)):            # target waves sized by the workspace budget
            cnt = min(wave, nt - a)
            part = (C.c_float * 4)()
            _lib.check(lib.orie_ensemble_sample(eng._handle, t0 + a, cnt, Nc, seed, C.c_void_p(bits.data_ptr()), s))
            out_r = C.c_void_p(0) if by_class else C.c_void_p(mine.data_ptr() + 8 * a)
            out_s = C.c_void_p(sums.data_ptr() + 24 * a) if by_class else C.c_void_p(0)
            if profile:      # untimed passes only: events around every kernel, synchronises
                _lib.check(lib.orie_reward_profile(eng._handle, t0 + a, cnt, C.c_void_p(bits.data_ptr()), Nc,
                                                   C.c_void_p(ws.data_ptr()), ws.numel(), out_r, out_s, 0, s, part))
                ms = [x + float(y) for x, y in zip(ms, part)]
            elif by_class:
                _lib.check(lib.orie_reward_sums(eng._handle, t0 + a, cnt, C.c_void_p(bits.data_ptr()), Nc,
                                                C.c_void_p(ws.data_ptr()), ws.numel(), out_s, 0, s))
            else:
                _lib.check(lib.orie_reward(eng._handle, t0 + a, cnt, C.c_void_p(bits.data_ptr()), Nc,
                                           C.c_void_p(ws.data_ptr()), ws.numel(), out_r, C.c_void_p(0), s))
        if by_class:
            dist.all_reduce(sums)                       # 3 doubles per target: AP sums are additive over classes
            gathered[:M].copy_(rewards_from_sums(sums, T, Nc))
        elif world > 1:
            dist.all_gather_into_tensor(gathered, mine)
        else:
            gathered.copy_(mine)
        e2.record()
        e2.synchronize()
        if profile:
            for k, v in zip(("label_walk_ms", "walk_ms", "ap_ms", "finalize_ms"), ms):
                kernel_ms[k].append(float(v))
        elif record:
            phase_ms["match_index_ms"].append(e0.elapsed_time(e1))
            phase_ms["reward_ms"].append(e1.elapsed_time(e2))
        info = eng.info
        eng.close()
        return e0.elapsed_time(e2), info

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                      # sampled from the warm-up to the end of the e2e runs
    for w in range(args.warmup):
        flush.fill_(w)
        step(1000 + w, False)
    barrier()
    launches0 = lib.orie_launch_count()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    info = None
    for k in range(args.steps):
        flush.fill_(k)                      # L2 flush, not timed
        torch.cuda.synchronize()
        ms, info = step(2000 + k, True)
        dev_ms += ms
    barrier()
    wall = time.perf_counter() - wall0
    launches = lib.orie_launch_count() - launches0
    for k in range(min(args.steps, 5)):     # per-kernel durations for the roofline: same step, events around each kernel, untimed
        flush.fill_(k)
        torch.cuda.synchronize()
        step(2000 + k, False, profile=True)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = M * args.steps / (dev_ms / 1e3)

    # ---- e2e: pinned host buffers -> rewards on the host, through the public API
    e2e_times = []
    e2e_warm = max(3, args.warmup)           # the first passes pay for pinned staging / allocator pools, like any warm-up
    for k in range(e2e_warm + max(2, min(args.steps, 10))):
        barrier()
        t_start = time.perf_counter()
        eng = Engine(hp, iouv=iouv, device=dev)
        if by_class:
            sums = eng.orie_sums_device(N, seed=3000 + k, total_images=M).clone()
            eng.stream.synchronize()
            dist.all_reduce(sums)
            host = rewards_from_sums(sums, T, min(N, M - 1)).cpu()
        else:
            mine = torch.zeros(per, dtype=torch.float64, device=dev)
            if nt > 0:
                mine[:nt] = eng.orie_device(N, seed=3000 + k, t0=t0, nt=nt)
            if world > 1:
                dist.all_gather_into_tensor(gathered, mine)
                host = gathered[:M].cpu()
            else:
                host = mine[:M].cpu()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t_start
        eng.close()
        if k >= e2e_warm:
            e2e_times.append(dt)
    t = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = M * len(e2e_times) / float(t.item())
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events around each kernel, averaged over the timed steps)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    per_target = algorithmic_bytes(pk, min(N, M - 1))      # pk = this rank's share of the records when classes are sharded
    alg_bytes = float(per_target[t0:t0 + nt].sum())          # bytes one launch on this rank accounts for
    means = {k: (sum(v) / len(v) if v else 0.0) for k, v in kernel_ms.items()}
    dom = max(("walk_ms", "ap_ms"), key=lambda k: means[k])

    def profiled_traffic(kernel_prefix):
        """dram__bytes_read.sum + dram__bytes_write.sum of the kernel from the committed `ncu --set full` capture of
        this same command (profiles/r01_ncu_raw_selected.csv), per launch; None if there is no capture."""
        import csv
        path = os.path.join(ROOT, "profiles", "r01_ncu_raw_selected.csv")
        if not os.path.exists(path) or args.workload != "coco5000" or world != 1:
            return None
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if kernel_prefix in r[0]:
                tot = 0.0
                for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = hdr.index(col)
                    tot += float(r[i]) * scale.get(units[i], 1.0)
                return tot
        return None
    achieved = alg_bytes / (means[dom] / 1e3) / 1e9 if means[dom] > 0 else 0.0
    reward_phase_ms = sum(means.values())
    roofline = {"bound": "hbm", "kernel": {"walk_ms": "walk_kernel<true>", "ap_ms": "ap_kernel"}[dom],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": profiled_traffic({"walk_ms": "walk_kernel<1", "ap_ms": "ap_kernel"}[dom]),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": means,
                "reward_phase": {"ms": reward_phase_ms, "achieved": alg_bytes / (reward_phase_ms / 1e3) / 1e9 if reward_phase_ms else 0.0},
                "note": "algorithmic bytes = the records upstream gathers per target (SURVEY 8d); the engine never moves them "
                        "(dataset and index are L2-resident, ensembles are membership tests, the AP sweep stops when the two "
                        "variants can no longer differ), so frac can exceed 1 - see DESIGN.md section 4 and the ncu traffic"}

    # ---- CPU baseline (rank 0, one thread, bounded sample) + live parity check on that sample
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:          # contract: CPU baseline on rank 0 at N = 1 only
        from oracle import orie_oracle as O
        eng = Engine(dp, iouv=iouv)
        wtp, stp, _, _ = eng.tp_flags()
        wd, sd, lc = cpu_cache_from_packed(pk, iouv, wtp, stp)
        S = min(args.cpu_sample, M)
        targets = np.linspace(0, M - 1, S).astype(np.int64)
        bits = eng.sample_bits(N, seed=77)
        member = ((bits[targets][:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(S, -1)[:, :M].astype(bool)
        got = eng.orie(N, seed=77)[targets]
        eng.close()
        tc = time.perf_counter()
        want = np.array([O.orie_one(int(i), wd, sd, lc, np.nonzero(member[r])[0])[0] for r, i in enumerate(targets)])
        tc = time.perf_counter() - tc
        want = np.where(np.isnan(want), 0, want)
        parity = float(np.abs(got - want).max())
        cpu = {"value": S / tc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{S} evenly spaced targets of the same workload, one thread, oracle/orie_oracle.py "
                         f"(reward phase only, TP cache prebuilt) on a host with {os.cpu_count()} cpus"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "images": M, "classes": pk_all.num_classes, "num_ensemble": min(N, M - 1),
                   "iou_thresholds": T, "weak_dets": int(len(pk_all.w_cls)), "strong_dets": int(len(pk_all.s_cls)),
                   "labels": int(len(pk_all.l_cls)), "parallelism": (f"classes sharded over {world} gpus (every rank: all targets, its classes), one all-reduce of 3 doubles "
                                   f"per target" if by_class else f"targets sharded over {world} gpu(s), index replicated, one all-gather"),
                   "step": "TP matching (2 detectors) + index build + ensemble draw + membership walk + AP + collective",
                   "l2": "flushed between steps (256 MiB write, not timed)", "ensembles": "device-side Philox draw, seed per step",
                   "index": {k: info[k] for k in ("slots", "segments", "events", "seg_chunks", "class_groups")},
                   "phase_ms": {k: sum(v) / len(v) for k, v in phase_ms.items()}, "wall_s_timed_region": wall},
        "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pk.nbytes()), "d2h_bytes_per_step": int(M * 8),
                "note": "bytes per rank" if world > 1 else "",
                "ms_per_step": 1e3 * sum(e2e_times) / len(e2e_times), "timer": "wall clock around Engine(pinned host) + orie + .cpu()"},
        "roofline": roofline, "cpu_baseline": cpu, "parity_max_abs_err_vs_oracle": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
