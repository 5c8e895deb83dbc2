"""CPU, world_size 2 and 4 over gloo: the multi-GPU host logic of the engine itself — ``shard_plan`` /
``shard_of_rank`` (class groups x target blocks), ``class_shard`` (the rows a rank keeps) and ``combine_sums`` (the one
all-reduce of zero-padded per-target AP sums + ``rewards_from_sums``) — run on CPU tensors.  What a rank's GPU would
compute (the AP sums of ITS classes for the targets of ITS block) is computed by the oracle on the rank's shard, so the
test fails if any part of the decomposition is wrong: classes assigned twice or not at all, target blocks that overlap
or leave a gap, the ensemble-size clamp taken from the shard instead of the dataset, the (N+1) multiplier."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _dataset(M):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_packed
    return make_packed(M=M, seed=77, zipf=0.7)[1]


def _oracle_sums(pk, iouv, em, t0, nt):
    """f64[nt, 3] (sum of weak APs, sum of strong APs, classes with ground truth) of targets [t0, t0+nt) on ``pk``."""
    from helpers import O, oracle_cache
    wd, sd, lc = oracle_cache(pk, iouv)
    out = np.zeros((nt, 3))
    for r in range(nt):
        _, wap, sap = O.orie_one(t0 + r, wd, sd, lc, em[t0 + r])
        out[r] = wap.sum(), sap.sum(), wap.shape[0]
    return out


def _worker(rank, world, port, M, N, shard, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orie_b200  # noqa: F401
    from helpers import O
    from orie_b200.engine import clamp_ensemble, class_shard, combine_sums, shard_of_rank
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pk = _dataset(M)
    iouv = O.IOU_05_095
    em = O.ensemble_matrix(M, N, 5)
    rc, rc_n, t0, nt = shard_of_rank(rank, M, world, shard)
    mine = class_shard(pk, rc, rc_n) if rc_n > 1 else pk
    sums = torch.from_numpy(_oracle_sums(mine, iouv, em, t0, nt))            # stand-in for orie_reward_sums on this rank's GPU
    reward = combine_sums(sums, t0, M, len(iouv), clamp_ensemble(M, N))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), reward.numpy())
    np.save(os.path.join(out_dir, f"p{rank}.npy"), np.array([rc, rc_n, t0, nt]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shard", [(2, "classes"), (2, "targets"), (4, "grid:2x2"), (2, "auto")])
def test_sharded_job_reassembles_the_rewards(tmp_path, world, shard):
    M, N = 70, 20
    mp.spawn(_worker, args=(world, _free_port(), M, N, shard, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import O, oracle_cache
    pk = _dataset(M)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    want = O.orie_all(wd, sd, lc, O.ensemble_matrix(M, N, 5))
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert got.shape == (M,) and np.abs(got - want).max() < 1e-11, (r, float(np.abs(got - want).max()))
    plans = np.stack([np.load(tmp_path / f"p{r}.npy") for r in range(world)])
    # every (class group, target) pair is owned exactly once
    owned = np.zeros((int(plans[0, 1]), M), dtype=int)
    for rc, _, t0, nt in plans:
        owned[rc, t0:t0 + nt] += 1
    assert (owned == 1).all()


def test_shard_plan_and_class_partition():
    sys.path.insert(0, ROOT)
    import orie_b200  # noqa: F401
    from orie_b200.engine import class_partition, class_shard, shard_plan, shard_range
    assert shard_plan(5000, 8) == (8, 1) and shard_plan(5000, 1) == (1, 1) and shard_plan(50000, 8) == (8, 1)
    assert shard_plan(5000, 8, num_classes=4) == (4, 2) and shard_plan(5000, 6, num_classes=4) == (3, 2)
    assert shard_plan(50000, 8, "grid:4x2") == (4, 2) and shard_plan(5000, 8, "targets") == (1, 8)
    with pytest.raises(ValueError):
        shard_plan(5000, 8, "grid:3x2")
    pk = _dataset(60)
    for world in (2, 3, 8):
        owner = class_partition(pk, world)
        assert owner.min() >= 0 and owner.max() < world
        rows = sum(len(class_shard(pk, r, world).w_cls) + len(class_shard(pk, r, world).s_cls) + len(class_shard(pk, r, world).l_cls)
                   for r in range(world))
        assert rows == len(pk.w_cls) + len(pk.s_cls) + len(pk.l_cls)
        cover = np.zeros(60, dtype=int)
        for r in range(world):
            t0, nt = shard_range(60, r, world)
            assert t0 % 32 == 0 or nt == 0
            cover[t0:t0 + nt] += 1
        assert (cover == 1).all()
