#!/usr/bin/env python
"""Calculate the offloading reward value for each image in a dataset (B200 engine).

Drop-in for the reference's ``reward.py`` (same positionals, ``--method`` and
``--num-ensemble``; same ``orie{N}.npz`` / ``dcsb.npz`` output with keys
``reward`` and ``time``).  Additions, all defaulting to the reference's
behaviour: ``--method ori`` (alias of ``orie --num-ensemble 0``),
``--iou-thresholds`` (the reference hard-codes [0.5] and keeps 0.5:0.95 as a
commented line, lib/data.py:60-62), ``--seed`` and ``--ensembles``.

Multi-GPU: ``python -m torch.distributed.run --nproc-per-node N reward.py ...``
shards the classes over the ranks (one all-reduce of the per-target AP sums); rank 0 writes the file.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def getargs(argv=None):
    """Parse command line arguments."""
    args = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    args.add_argument('weak_dir', help="Directory to the weak detector output files.")
    args.add_argument('strong_dir', help="Directory to the strong detector output files.")
    args.add_argument('label_dir', help="Directory to the ground truth annotations.")
    args.add_argument('save_dir', help="Directory to save the computed offloading rewards.")
    args.add_argument('--method', type=str, default="orie", choices=['orie', 'ori', 'dcsb'],
                      help="Method used to compute the offloading reward ('ori' = 'orie' with --num-ensemble 0).")
    args.add_argument('--num-ensemble', type=int, default=1000,
                      help="Number of ensemble images when computing the offloading reward, only active when method "
                           "is 'orie', in which case setting num-ensemble to 0 yields ORI as the reward metric.")
    args.add_argument('--iou-thresholds', type=str, default="0.5",
                      help="'0.5' (mAP@0.5, the reference's shipped setting), '0.5:0.95' (ten COCO thresholds) or a "
                           "comma separated list (at most 16).")
    args.add_argument('--seed', type=int, default=None,
                      help="Seed of the ensemble draw (default: fresh entropy, like the unseeded reference).")
    args.add_argument('--ensembles', type=str, default="device", choices=['device', 'numpy'],
                      help="'device': counter-based draw on the GPU; 'numpy': regenerate the reference's "
                           "np.random.permutation draw on the host with seed+image index (parity runs).")
    args.add_argument('--shard', type=str, default="auto", choices=['auto', 'classes', 'targets'],
                      help="Multi-GPU decomposition under torchrun: 'classes' (one all-reduce of per-target AP sums), "
                           "'targets' (one all-gather of reward slices), 'auto' = by dataset size.")
    return args.parse_args(argv)


def main(opts):
    import orie_b200  # noqa: F401
    from orie_b200 import api
    reward, seconds, info = api.compute_rewards_from_dirs(
        opts.weak_dir, opts.strong_dir, opts.label_dir, method=opts.method, num_ensemble=opts.num_ensemble,
        iouv=opts.iou_thresholds, seed=opts.seed, ensembles=opts.ensembles, shard=opts.shard)
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"Program takes {seconds:.1f} seconds ({seconds / 60:.1f}m/{seconds / 3600:.2f}h).")
        path = api.save_rewards(opts.save_dir, opts.method, opts.num_ensemble, reward, seconds)
        print(f"Saved {len(reward)} rewards to {path} (loading {info['load_s']:.1f}s, "
              f"upload + matching {info.get('match_index_s', 0.0):.2f}s).")
    return


if __name__ == '__main__':
    main(getargs())
