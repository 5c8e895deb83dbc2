"""orie-b200: B200-native engine for the ORIE / ORI / DCSB offloading-reward
path of qiujiaming315/edgeml-object-detection (reward.py + lib/metrics.py +
the loaders of lib/data.py).  See DESIGN.md."""
from . import synth, data  # noqa: F401

__version__ = "0.1.0"
