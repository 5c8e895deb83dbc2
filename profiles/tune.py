"""Developer tool (not part of the product path): A/B kernel timings inside one process."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import orie_b200  # noqa
from orie_b200.engine import DevicePacked, Engine, HostPacked
sys.argv = [sys.argv[0]] + sys.argv[1:]
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "coco5000"
segs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,16,32,64,128".split(","))]
ds, pk, method, N, iouv = bench.dataset(workload)
dev = torch.device("cuda:0")
hp = HostPacked(pk)
dp = DevicePacked(hp, dev)
torch.cuda.synchronize()
waves = os.environ.get("TUNE_WAVES", "").split(",") if os.environ.get("TUNE_WAVES") else [None]
for sc in [(a, w) for a in segs for w in waves]:
    sc, wv = sc
    eng = Engine(dp, iouv=iouv, seg_chunks=sc, tuning=dict(walk_waves=float(wv)) if wv is not None else None)
    rows = []
    for rep in range(6):
        rows.append(eng.profile_reward(N, seed=rep))
    med = {k: float(np.median([r[k] for r in rows[1:]])) for k in rows[0]}
    print("seg_chunks", sc, "waves", wv, eng.info["segments"], {k: round(v, 3) for k, v in med.items()}, flush=True)
    eng.close()

# e2e breakdown
def sync():
    torch.cuda.synchronize()
for rep in range(6):
    t = [time.perf_counter()]
    d = DevicePacked(hp, dev); sync(); t.append(time.perf_counter())
    eng = Engine(d, iouv=iouv); sync(); t.append(time.perf_counter())
    r = eng.orie_device(N, seed=rep); sync(); t.append(time.perf_counter())
    h = r.cpu(); sync(); t.append(time.perf_counter())
    eng.close(); sync(); t.append(time.perf_counter())
    print("e2e ms: h2d %.2f  match+index %.2f  reward %.2f  d2h %.2f  close %.2f" % tuple(1e3 * (b - a) for a, b in zip(t, t[1:])), flush=True)

# pipelined (pinned host -> Engine) breakdown, as bench.py's e2e does it
for rep in range(6):
    sync()
    t = [time.perf_counter()]
    eng = Engine(hp, iouv=iouv, device=dev); t.append(time.perf_counter())
    sync(); t.append(time.perf_counter())
    r = eng.orie_device(N, seed=rep); t.append(time.perf_counter())
    h = r.cpu(); sync(); t.append(time.perf_counter())
    eng.close()
    print("pipelined ms: Engine() returns %.2f  +sync %.2f  orie enqueue %.2f  cpu() %.2f   total %.2f" %
          (tuple(1e3 * (b - a) for a, b in zip(t, t[1:])) + (1e3 * (t[-1] - t[0]),)), flush=True)
