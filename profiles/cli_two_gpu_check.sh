set -e
python - <<'PY'
import sys; sys.path.insert(0, '.')
import orie_b200
from orie_b200 import synth
ds = synth.make("smoke500", num_images=200, seed=3)
synth.write_dirs(ds, "/tmp/cli_ds")
PY
D=/tmp/cli_ds
python reward.py $D/weak $D/strong $D/labels /tmp/out1 --num-ensemble 50 --seed 9 --iou-thresholds 0.5:0.95 > /dev/null
for sh in targets classes; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 reward.py $D/weak $D/strong $D/labels /tmp/out_$sh --num-ensemble 50 --seed 9 --iou-thresholds 0.5:0.95 --shard $sh > /dev/null 2>&1
python -c "
import numpy as np
a=np.load('/tmp/out1/orie50.npz')['reward']; b=np.load('/tmp/out_$sh/orie50.npz')['reward']
print('$sh', 'cli 1 vs 2 gpus max diff', float(np.abs(a-b).max()), len(a))"
done
