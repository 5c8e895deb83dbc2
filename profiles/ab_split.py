"""Developer tool: does cutting a reward call into target pieces on two streams pay (AP of one piece next to the walk of
the other)?      python profiles/ab_split.py [workload]"""
import ctypes as C, json, os, sys
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
lib, M = eng.lib, pk.num_images
N = min(N, M - 1)
torch.cuda.synchronize()
info = eng.info
bits = torch.empty((M, eng.ens_words), dtype=torch.int32, device=dev)
main = torch.cuda.current_stream()
side = torch.cuda.Stream(device=dev)
s_main, s_side = C.c_void_p(main.cuda_stream), C.c_void_p(side.cuda_stream)
_lib.check(lib.orie_ensemble_sample(eng._handle, 0, M, N, 7, C.c_void_p(bits.data_ptr()), s_main))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty(M, dtype=torch.float64, device=dev)
ref = None


def run(cuts):
    """cuts: list of (t0, nt, stream index)"""
    wss = [torch.empty(lib.orie_reward_workspace_bytes(eng._handle, nt), dtype=torch.uint8, device=dev) for _, nt, _ in cuts]
    ms = []
    for rep in range(7):
        flush.fill_(rep); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        side.wait_stream(main)
        for (t0, nt, si), ws in zip(cuts, wss):
            st = s_main if si == 0 else s_side
            _lib.check(lib.orie_reward(eng._handle, t0, nt, C.c_void_p(bits.data_ptr() + 4 * eng.ens_words * t0), N,
                                       C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(out.data_ptr() + 8 * t0), C.c_void_p(0), st))
        main.wait_stream(side)
        b.record(main); b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms[2:])), out.cpu().numpy().copy()


def pieces(k, streams):
    per = ((M + k - 1) // k + 63) // 64 * 64
    cuts, t0, i = [], 0, 0
    while t0 < M:
        cuts.append((t0, min(per, M - t0), i % streams))
        t0 += per; i += 1
    return cuts


for name, cuts in [("one call", pieces(1, 1)), ("2 pieces, 2 streams", pieces(2, 2)), ("4 pieces, 2 streams", pieces(4, 2)),
                   ("2 pieces, 1 stream", pieces(2, 1)), ("8 pieces, 2 streams", pieces(8, 2))]:
    t, r = run(cuts)
    if ref is None:
        ref = r
    print(json.dumps({"variant": name, "reward_pass_ms": round(t, 4), "max_diff_vs_one_call": float(np.abs(r - ref).max())}), flush=True)
