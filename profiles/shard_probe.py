"""Developer tool: what ONE rank of an R-way class-sharded run executes, on one GPU (for ncu launch lists).
    python profiles/shard_probe.py [workload] [R class groups] [replays] [target blocks = 1]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import orie_b200  # noqa
from orie_b200.engine import DevicePacked, HostPacked, ReplayJob, class_shard
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "coco5000"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
Rt = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ds, pk, method, N, iouv = bench.dataset(workload)
M = pk.num_images
sh = class_shard(pk, 0, R) if R > 1 else pk
dp = DevicePacked(HostPacked(sh), "cuda")
from orie_b200.engine import shard_range
t0, nt = shard_range(M, 0, Rt)
job = ReplayJob(dp, iouv=iouv, num_ensemble=N, t0=t0, nt=nt, sums=R * Rt > 1, total_images=M)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for k in range(reps):
    flush.fill_(k); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); job.run(100 + k); b.record(); b.synchronize()
    ms.append(a.elapsed_time(b))
print(f"{workload} rank 0 of {R} class groups x {Rt} target blocks ({nt} targets): rows {len(sh.w_cls)}+{len(sh.s_cls)}, classes {sh.num_classes}, replay ms {np.round(ms, 4).tolist()}, "
      f"{job.launches_per_replay} kernels per replay")
