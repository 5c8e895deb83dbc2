"""GPU: the two consumers of the reward vector against golden vectors frozen from the LIVE reference
(oracle/gen_golden_consumers.py): baseline.fit_dcsb (baseline.py:67-152) and the rank normalisation of
regression.py:439-441, both through the C ABI (orie_dcsb_fit, orie_rank_normalize)."""
import os

import numpy as np
import pytest

from orie_b200 import api, data
from orie_b200.synth import Rows

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["dcsb_fit_coco", "dcsb_fit_voc"])
def test_dcsb_fit_reproduces_the_reference(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    pk = data.pack(Rows(z["l_off"], z["l_rows"]), Rows(z["w_off"], z["w_rows"]), Rows(z["s_off"], z["s_rows"]))
    fold = z["fold"]
    for k in range(int(fold.max()) + 1):
        val = fold == k
        got = api.fit_dcsb(pk, z["reward"], val)
        conf_t, num_t, area_t = z[f"model{k}"]
        assert got["conf_thresh"] == conf_t and got["num_thresh"] == int(num_t) and got["area_thresh"] == area_t     # bit-exact
        assert np.array_equal(got["train_est"], z[f"train_est{k}"]) and np.array_equal(got["val_est"], z[f"val_est{k}"])


def test_dcsb_fit_without_validation_rows_and_bad_arguments():
    z = np.load(os.path.join(GOLDEN, "dcsb_fit_coco.npz"))
    pk = data.pack(Rows(z["l_off"], z["l_rows"]), Rows(z["w_off"], z["w_rows"]), Rows(z["s_off"], z["s_rows"]))
    got = api.fit_dcsb(pk, z["reward"])
    assert len(got["val_est"]) == 0 and len(got["train_est"]) == pk.num_images and set(np.unique(got["est"])) <= {0, 1}
    # the decisions follow the fitted rule
    for i in range(pk.num_images):
        a, b = pk.w_off[i], pk.w_off[i + 1]
        conf = pk.w_conf[a:b]
        sel = conf > got["conf_thresh"]
        area = (pk.w_box[a:b, 2] - pk.w_box[a:b, 0]) * (pk.w_box[a:b, 3] - pk.w_box[a:b, 1])
        num, amin = int(sel.sum()), (area[sel].min() if sel.any() else 0.0)
        want = int(num != int((conf > 0.5).sum()) and (num > got["num_thresh"] or amin < got["area_thresh"]))
        assert got["est"][i] == want
    with pytest.raises(ValueError):
        api.fit_dcsb(pk, z["reward"][:-1])


def test_rank_normalize_reproduces_the_reference_lines():
    z = np.load(os.path.join(GOLDEN, "rank_norm.npz"))
    for k in range(int(z["cases"])):
        got = api.rank_normalize(z[f"reward{k}"], z[f"val{k}"])
        assert np.array_equal(got, z[f"want{k}"])
