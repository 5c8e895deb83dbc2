"""GPU: rank normalisation of rewards (orie_rank_normalize) against the reference's formula, regression.py:439-441."""
import numpy as np
import pytest

from orie_b200 import api

pytestmark = pytest.mark.gpu


def _upstream(reward, val_mask):
    """regression.py:439-441 verbatim in effect; ties broken in row order (kind='stable')."""
    train, val = reward[~val_mask], reward[val_mask]
    out = np.empty(len(reward))
    out[val_mask] = np.array([np.sum(train <= x) / len(train) for x in val]) if len(val) else []
    out[~val_mask] = (np.argsort(np.argsort(train, kind="stable"), kind="stable") + 1) / len(train)
    return out


@pytest.mark.parametrize("M,frac", [(1, 0.0), (7, 0.3), (500, 0.2), (5000, 0.2), (40000, 0.5)])
def test_rank_normalize_matches_the_reference_formula(M, frac):
    rng = np.random.default_rng(M)
    reward = rng.normal(size=M) * 10.0 ** rng.integers(-8, 3, size=M)        # tie-free, both signs, many magnitudes
    val_mask = rng.random(M) < frac
    if val_mask.all():
        val_mask[0] = False
    assert np.array_equal(api.rank_normalize(reward, val_mask), _upstream(reward, val_mask))
    assert np.array_equal(api.rank_normalize(reward), _upstream(reward, np.zeros(M, dtype=bool)))


def test_rank_normalize_with_ties_and_zeros():
    rng = np.random.default_rng(3)
    reward = np.round(rng.normal(size=3000), 1)            # heavy ties
    reward[::7] = 0.0
    reward[3::14] = -0.0                                    # numpy compares -0.0 == 0.0
    val_mask = rng.random(3000) < 0.25
    got = api.rank_normalize(reward, val_mask)
    assert np.array_equal(got, _upstream(reward, val_mask))
    # what does not depend on the tie order at all: validation rows, and the multiset of train ranks
    n = int((~val_mask).sum())
    assert np.array_equal(np.sort(got[~val_mask]), (np.arange(n) + 1) / n)
