"""
The big-brain restatement (oracle/lev_oracle.py:big_brain) against fixtures produced
by the unmodified reference coin_big_brain_lev / dice_big_brain_lev
(tests/golden/gen_golden_more.py:gen_bigbrain).
"""
import numpy as np
import pytest

import golden_io
from oracle import lev_oracle as lo
from test_oracle_lev import assert_stats_close


def oracle_data(case):
    oc = golden_io.draw_outcomes(case)
    returns = (case["up_r"], case["down_r"]) + ((case["mid_r"],) if case["kind"] == "dice" else ())
    return lo.big_brain(case["kind"], oc, case["top"], case["v0"], returns, golden_io.bigbrain_lev_factor(case),
                        lo.param_range(*case["stop"]), lo.param_range(*case["roll"]))


def assert_bigbrain_close(data, want, exact_medians=True):
    """data / want: [R,S,26,T].  Medians (rows 9..11, 21..23) and the grid rows bit-exact."""
    assert data.shape == want.shape
    assert np.array_equal(data[:, :, 24:26], want[:, :, 24:26])
    for base in (0, 12):
        med_g, med_w = data[:, :, base + 9:base + 12], want[:, :, base + 9:base + 12]
        if exact_medians:
            assert np.array_equal(med_g.view(np.uint32), med_w.view(np.uint32)), \
                (base, np.argwhere(med_g != med_w)[:5])
        else:
            np.testing.assert_allclose(med_g, med_w, rtol=1e-5)
        assert_stats_close(data[:, :, base:base + 9], want[:, :, base:base + 9], rtol=3e-5, noise=2e-6)


@pytest.mark.parametrize("case", golden_io.BIGBRAIN_CASES, ids=lambda c: c["name"])
def test_big_brain_matches_reference(case):
    gold = golden_io.load("bigbrain_" + case["name"])
    data = oracle_data(case)[:, :, :, gold["cols"]]
    assert_bigbrain_close(data, gold["data"])


def test_galaxy_brain_matches_reference():
    g = golden_io.GALAXY_GRID
    want = golden_io.load("galaxy_brain")["data"]
    got = lo.galaxy_brain(lo.param_range(*g[0:3]), lo.param_range(*g[3:6]), lo.param_range(*g[6:9]))
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
