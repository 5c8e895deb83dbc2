import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import orie_b200
from orie_b200.engine import DevicePacked, Engine, HostPacked
import bench
ds, pk, N, iouv = bench.dataset("coco5000")
dev = torch.device("cuda:0")
hp = HostPacked(pk); dp = DevicePacked(hp, dev)
def one(seed):
    eng = Engine(dp, iouv=iouv)
    r = eng.orie_device(N, seed=seed)
    torch.cuda.synchronize()
    eng.close()
for i in range(5): one(i)
t=time.perf_counter()
for i in range(50): one(i)
print("wall per job ms", (time.perf_counter()-t)/50*1e3)
def init_only():
    eng = Engine(dp, iouv=iouv); eng.close()
t=time.perf_counter()
for i in range(50): init_only()
print("Engine() per call ms", (time.perf_counter()-t)/50*1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(50): one(i)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
