"""Developer tool (not part of the product path): A/B timings of kernel variants inside one process.
    python profiles/ab.py [workload] ['[{"walk_single": 1}, {"walk_waves": 3.0}]' | seg_chunks list "0,64"]
Prints per-kernel medians (orie_reward_profile) for tuning variants and the match+index phase time."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import orie_b200  # noqa
from orie_b200.engine import DevicePacked, Engine, HostPacked
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "coco5000"
ds, pk, method, N, iouv = bench.dataset(workload)
dev = torch.device("cuda:0")
dp = DevicePacked(HostPacked(pk), dev)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
arg = sys.argv[2] if len(sys.argv) > 2 else "0"
variants = json.loads(arg) if arg.lstrip().startswith("[") else [dict(seg_chunks=int(x)) for x in arg.split(",")]
ref = None
for tv in variants:
    rows, idx = [], []
    for rep in range(6):
        flush.fill_(rep); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng = Engine(dp, iouv=iouv, tuning=tv)
        b.record(); b.synchronize()
        idx.append(a.elapsed_time(b))
        rows.append(eng.profile_reward(N, seed=rep))
        if rep == 5:
            r = eng.orie(N, seed=5)
            if ref is None:
                ref = r
            err = float(np.abs(r - ref).max())
        eng.close()
    med = {k: round(float(np.median([r[k] for r in rows[1:]])), 4) for k in rows[0]}
    import time as _t
    print(json.dumps({"tuning": tv, "match_index_ms": round(float(np.median(idx[1:])), 4), **med, "max_diff_vs_default": err}), flush=True)
