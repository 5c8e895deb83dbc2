"""Developer tool: distribution of the AP kernel's per-(target, class, threshold) sweep depths (orie_reward_depths)."""
import ctypes as C, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import orie_b200  # noqa
from orie_b200 import _lib
from orie_b200.engine import DevicePacked, Engine, HostPacked
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "coco5000"
ds, pk, method, N, iouv = bench.dataset(workload)
dev = torch.device("cuda:0")
eng = Engine(DevicePacked(HostPacked(pk), dev), iouv=iouv)
M, Cn, T = pk.num_images, pk.num_classes, len(iouv)
N = min(N, M - 1)
nt = min(M, 1024)
ws = eng._workspace(eng.workspace_bytes(nt))
bits = torch.empty((nt, eng.ens_words), dtype=torch.int32, device=dev)
rw = torch.empty(nt, dtype=torch.float64, device=dev)
dep = torch.zeros((nt, Cn, T, 2), dtype=torch.int32, device=dev)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
lib = eng.lib
_lib.check(lib.orie_ensemble_sample(eng._handle, 0, nt, N, 5, C.c_void_p(bits.data_ptr()), s))
_lib.check(lib.orie_reward_depths(eng._handle, 0, nt, C.c_void_p(bits.data_ptr()), N, C.c_void_p(ws.data_ptr()), ws.numel(),
                                  C.c_void_p(rw.data_ptr()), C.c_void_p(dep.data_ptr()), s))
d = dep.cpu().numpy().astype(np.int64)
main, tail = d[..., 0], d[..., 1]
tot = main + tail
act = tot.sum(axis=2) > 0                       # (target, class) integrated
print("targets", nt, "active classes per target: mean %.1f" % act.sum(1).mean())
q = [50, 75, 90, 95, 99, 100]
print("trips per (class,thr) item, active only: main", np.percentile(main[act], q).tolist(), "tail", np.percentile(tail[act], q).tolist())
print("mean main %.1f tail %.1f" % (main[act].mean(), tail[act].mean()))
print("by threshold: mean main", main[act].mean(0).round(1).tolist())
print("by threshold: mean tail", tail[act].mean(0).round(1).tolist())
cls_depth = tot.max(axis=2)                     # a class's lanes run as long as its deepest threshold
print("per-class max over thresholds (active): pct", np.percentile(cls_depth[act], q).tolist(), "mean %.1f" % cls_depth[act].mean())
print("sum of trips over all lanes / (sum over classes of max-lane trips * T) = lane efficiency inside a class: %.3f"
      % (tot[act].sum() / (cls_depth[act].sum() * T)))
# what a warp of 3 classes pays today: max over its lanes; fixed cls_order groups vs ideal packing by depth
order = np.argsort(-np.bincount(pk.w_cls, minlength=Cn), kind="stable")
cpw = 32 // T
pad = (-Cn) % cpw
g = np.concatenate([cls_depth[:, order], np.zeros((nt, pad), dtype=np.int64)], axis=1).reshape(nt, -1, cpw)
print("warp trips (fixed groups): total %.3e" % g.max(axis=2).sum())
srt = -np.sort(-cls_depth, axis=1)
g2 = np.concatenate([srt, np.zeros((nt, pad), dtype=np.int64)], axis=1).reshape(nt, -1, cpw)
print("warp trips (classes packed by depth per target): total %.3e" % g2.max(axis=2).sum())
print("lane trips total %.3e  -> ideal warp trips at 30 lanes %.3e" % (tot.sum(), tot.sum() / 30))
