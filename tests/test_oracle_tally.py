"""
The count-tuple route on the CPU (oracle/tally_oracle.py) against the per-investor oracle
(pinned to the reference's goldens): the algorithm the GPU's tally path implements gives the
same 12 statistics - order statistics bit for bit, fp64 moments to 1e-12 - including ties
across the top-K boundary, infinities, zeros and one-investor groups.
"""
import numpy as np
import pytest

import golden_io
from oracle import lev_oracle as lo
from oracle import tally_oracle as to

DISCRETE = [c for c in golden_io.LEV_CASES if c["kind"] != "gbm"]


def factors_of(case):
    lev = lo.lev_grid(*case["grid"], case["up_r"], case["down_r"])
    if case["kind"] == "coin":
        return lo.coin_factors(lev, case["up_r"], case["down_r"])
    if case["kind"] == "dice":
        return lo.dice_factors(lev, case["up_r"], case["down_r"], case["mid_r"])
    return lo.dice_sh_factors(lev, case["up_r"], case["down_r"], case["mid_r"], *case["sh"])


def per_investor(oc, f, top, v0):
    with np.errstate(over="ignore", invalid="ignore"):
        w = np.exp(lo.log_wealth_discrete(oc, f, v0)).astype(np.float32)
        return np.stack([lo.summary_stats(w[g], top) for g in range(w.shape[0])])


def assert_same(got, want):
    assert np.array_equal(np.isnan(got), np.isnan(want)), (got, want)
    ok = ~np.isnan(want)
    assert np.array_equal(got[:, 9:12][ok[:, 9:12]], want[:, 9:12][ok[:, 9:12]])          # medians: the same fp32 value
    fin = np.isfinite(want)
    assert np.array_equal(got[ok & ~fin], want[ok & ~fin])
    scale = np.where(np.isfinite(want[:, 0:1]), np.abs(want[:, 0:1]), 0.0) + 1e-300
    with np.errstate(invalid="ignore"):
        err = np.abs(got - want)
    assert ((err <= 1e-12 * np.maximum(np.abs(want), scale)) | ~fin).all()


@pytest.mark.parametrize("case", DISCRETE, ids=lambda c: c["name"])
def test_count_tuple_statistics_equal_the_per_investor_ones(case):
    oc = golden_io.draw_outcomes(case)
    f = factors_of(case)
    tuples, counts = to.bins_of(oc, f.shape[1])
    assert counts.sum() == oc.shape[0] and len(tuples) <= oc.shape[0]
    assert_same(to.final_stats(oc, f, case["top"], case["v0"]), per_investor(oc, f, case["top"], case["v0"]))


@pytest.mark.parametrize("top", [1, 2, 7, 39, 40, 41, 199])
def test_ties_across_the_top_boundary_and_degenerate_rows(top):
    """Few distinct tuples (heavy ties around rank K), a leverage whose wealth is 0 or inf, K = 1 and K = n - 1."""
    rs = np.random.RandomState(top)
    n, h = 200, 12
    oc = rs.choice(3, size=(n, h), p=[0.2, 0.2, 0.6]).astype(np.uint8)
    oc[:40] = oc[0]                                   # forty investors share one tuple
    f = np.float32([[1.5, 0.5, 1.05], [3.0e30, 0.0, 1.0], [1.0, 1.0, 1.0], [2.0, 0.0, 1.0]])
    assert_same(to.final_stats(oc, f, top, 100.0), per_investor(oc, f, top, 100.0))
