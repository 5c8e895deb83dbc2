"""CPU, world_size 2 over gloo: the multi-GPU host logic (contiguous target
shards on 32-boundaries, zero-padded slices, one all-gather) without a GPU."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, M, out_dir):
    sys.path.insert(0, ROOT)
    import orie_b200  # noqa: F401
    from orie_b200.engine import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t0, nt = shard_range(M, rank, world)
    per = shard_range(M, 0, world)[1]
    mine = torch.zeros(per, dtype=torch.float64)
    mine[:nt] = torch.arange(t0, t0 + nt, dtype=torch.float64) * 0.5 + 1.0     # stand-in for this rank's rewards
    parts = [torch.empty(per, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, mine)
    full = torch.cat(parts)[:M].numpy()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    dist.destroy_process_group()


def _worker_classes(rank, world, port, M, out_dir):
    sys.path.insert(0, ROOT)
    import orie_b200  # noqa: F401
    from orie_b200.engine import rewards_from_sums
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    sums = torch.rand((M, 3), dtype=torch.float64, generator=g)          # stand-in for this rank's class-shard sums
    sums[:, 2] = torch.randint(0, 4, (M,), generator=g).double()
    np.save(os.path.join(out_dir, f"s{rank}.npy"), sums.numpy())
    dist.all_reduce(sums)
    np.save(os.path.join(out_dir, f"c{rank}.npy"), rewards_from_sums(sums, 10, 7).numpy())
    dist.destroy_process_group()


def test_two_rank_class_shard_all_reduce(tmp_path):
    M = 90
    mp.spawn(_worker_classes, args=(2, _free_port(), M, str(tmp_path)), nprocs=2, join=True)
    total = np.load(tmp_path / "s0.npy") + np.load(tmp_path / "s1.npy")
    nc = total[:, 2]
    want = np.where(nc > 0, (total[:, 1] - total[:, 0]) / np.maximum(nc * 10, 1) * 8, 0.0)
    for r in range(2):
        assert np.allclose(np.load(tmp_path / f"c{r}.npy"), want, rtol=0, atol=1e-12)


def test_two_rank_shard_and_gather(tmp_path):
    for M in (70, 500):
        port = _free_port()
        mp.spawn(_worker, args=(2, port, M, str(tmp_path)), nprocs=2, join=True)
        want = np.arange(M) * 0.5 + 1.0
        for r in range(2):
            assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), want)
