#!/usr/bin/env python
"""Realised mAP of reward estimates versus the offloading ratio (B200 engine).

Drop-in for the reference's ``test.py`` (same positionals and ``--estimates``; writes ``SAVE_DIR/test_map.npy`` with
one row of 11 mAP values per estimate directory).  ``--iou-thresholds`` defaults to the reference's shipped [0.5].
"""
import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def getargs(argv=None):
    """Parse command line arguments."""
    args = argparse.ArgumentParser(description=__doc__)
    args.add_argument('weak_dir', help="Directory to the weak detector output files.")
    args.add_argument('strong_dir', help="Directory to the strong detector output files.")
    args.add_argument('label_dir', help="Directory to the ground truth annotations.")
    args.add_argument('split_path', help="Path to the dataset split (for cross validation).")
    args.add_argument('save_dir', help="Directory to save the achieved mAP.")
    args.add_argument('--estimates', nargs='+', type=str, help='Directories to the reward estimation file(s).')
    args.add_argument('--iou-thresholds', type=str, default="0.5", help="'0.5' (reference default), '0.5:0.95' or a list.")
    return args.parse_args(argv)


def main(opts):
    import numpy as np
    import orie_b200  # noqa: F401
    from orie_b200 import api, evaluate
    result = evaluate.test_map_from_dirs(opts.weak_dir, opts.strong_dir, opts.label_dir, opts.split_path, opts.estimates,
                                         iouv=api.parse_iou_thresholds(opts.iou_thresholds))
    Path(opts.save_dir).mkdir(parents=True, exist_ok=True)
    np.save(os.path.join(opts.save_dir, 'test_map.npy'), result)
    return


if __name__ == '__main__':
    main(getargs())
