"""
Pins oracle/lev_oracle.py against the fixtures written by the unmodified
reference (tests/golden/gen_golden.py).  CPU only.
"""
import numpy as np
import pytest

import golden_io
from oracle import lev_oracle as lo


def oracle_inputs(case):
    oc = golden_io.draw_outcomes(case)
    kind = case["kind"]
    if kind == "gbm":
        lev = lo.lev_grid(*case["grid"])
        return oc, lev, lev, True
    lev = lo.lev_grid(*case["grid"], case["up_r"], case["down_r"])
    if kind == "coin":
        f = lo.coin_factors(lev, case["up_r"], case["down_r"])
    elif kind == "dice":
        f = lo.dice_factors(lev, case["up_r"], case["down_r"], case["mid_r"])
    else:
        f = lo.dice_sh_factors(lev, case["up_r"], case["down_r"], case["mid_r"], *case["sh"])
    return oc, f, lev, False


def assert_stats_close(got, want, rtol=2e-5, noise=1e-6):
    """
    Rows 0..8 (mean x3, mad x3, std x3) within rtol, plus an absolute floor of
    `noise` x |group mean| on mad/std: torch subtracts an fp32-rounded mean, so a
    group of (nearly) equal values reports a mad/std of a few ulps of the mean
    instead of 0.  nan/inf patterns must be equal.  Layout [..., rows, T].
    """
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf])
    fin = np.isfinite(want)
    atol = np.zeros_like(want)
    nrow = want.shape[-2]
    for r in range(3, min(nrow, 9)):
        atol[..., r, :] = noise * np.abs(np.where(np.isfinite(want[..., r % 3, :]), want[..., r % 3, :], 0.0))
    err = np.abs(got - want) - atol
    bad = fin & (err > rtol * np.abs(want))
    assert not bad.any(), (np.argwhere(bad)[:5], got[bad][:5], want[bad][:5])


@pytest.mark.parametrize("case", golden_io.LEV_CASES, ids=lambda c: c["name"])
def test_smart_lev_matches_reference(case):
    oc, f, lev, gbm = oracle_inputs(case)
    gold = golden_io.load("lev_" + case["name"], oc)
    if case["n"] * case["h"] > 1_000_000:
        # big cases: chain for data_T, statistics only at the kept columns
        data_T, steps = (lo.chain_gbm if gbm else lo.chain_discrete)(oc, f, case["v0"], keep_steps=True)
        cols = gold["cols"]
        data = np.zeros((data_T.shape[0], 13, len(cols)), dtype=np.float32)
        with np.errstate(all="ignore"):
            for j, t in enumerate(cols):
                for g in range(data_T.shape[0]):
                    data[g, :12, j] = lo.summary_stats(steps[t][g], case["top"])
                    data[g, 12, j] = lev[g]
    else:
        data, data_T = lo.smart_lev(oc, f, lev, case["top"], case["v0"], gbm=gbm)
        data = data[:, :, gold["cols"]]

    if gbm:
        # expf differs between numpy and torch (Sleef) in the last ulp
        fin = np.isfinite(gold["data_T"]) & (gold["data_T"] > 0)
        assert np.array_equal(np.isfinite(data_T) & (data_T > 0), fin) or \
            (np.isfinite(data_T) != np.isfinite(gold["data_T"])).mean() < 1e-3
        both = fin & np.isfinite(data_T)
        rel = np.abs(data_T[both].astype(np.float64) - gold["data_T"][both]) / gold["data_T"][both]
        assert rel.max() <= 2e-5 * max(1.0, case["h"] / 300)
    else:
        assert np.array_equal(data_T.view(np.uint32), gold["data_T"].view(np.uint32)), "data_T must be bit-exact"
        # medians are order statistics of bit-exact values: bit-exact too
        assert np.array_equal(data[:, 9:12].view(np.uint32), gold["data"][:, 9:12].view(np.uint32))
    # leverage row
    assert np.array_equal(data[:, 12], gold["data"][:, 12])
    if not gbm:
        assert_stats_close(data[:, :9], gold["data"][:, :9])
    elif case["name"] != "gbm_overflow":
        assert_stats_close(data[:, :12], gold["data"][:, :12], rtol=5e-5)
    else:
        # deep in the denormal range the fp32 chain has lost its relative accuracy:
        # compare where the reference is still normal, and the zero pattern elsewhere
        normal = np.abs(gold["data"][:, :12]) > 1e-30
        assert_stats_close(np.where(normal, data[:, :12], 0), np.where(normal, gold["data"][:, :12], 0), rtol=2e-4)
        assert np.array_equal(gold["data"][:, 9:12] == 0, data[:, 9:12] == 0)


def test_param_range_truth_table():
    """SURVEY.md App. A (measured on the reference)."""
    assert len(lo.param_range(0.05, 1.00, 0.05)) == 20
    assert len(lo.param_range(0.10, 1.00, 0.10)) == 10
    r = lo.param_range(0.05, 0.95, 0.05)
    assert len(r) == 18 and abs(r[-1] - 0.90) < 1e-12
    assert len(lo.param_range(0.70, 0.95, 0.05)) == 6
    assert lo.param_range(0.10, 0.10, 0.10) == [0.1]
    assert lo.param_range(0.0, 0.0, 0.1) == [0.0]
    r = lo.param_range(0.73, 1.00, 0.03)
    assert len(r) == 10 and r[-1] == 0.9999999999999998
    r = lo.param_range(-1.0, 1.0, 0.2)
    assert len(r) == 10 and 0 not in r
    assert len(lo.param_range(0.7, 0.8, 0.1)) == 3
    assert len(lo.param_range(0.2, 0.8, 0.001)) == 601


def test_percentile_type8_matches_numpy():
    rs = np.random.RandomState(0)
    v = rs.lognormal(size=1001).astype(np.float32)
    for q in (0.05, 0.5, 0.95):
        want = np.percentile(v.astype(np.float64), q * 100, method="median_unbiased")
        assert abs(lo.percentile_type8(v, q) - want) <= 1e-12 * abs(want)


def test_packed_code_layout_round_trips():
    """Engine format with outcome_bits = 2: four codes per byte, little end first, zero pad bits."""
    rs = np.random.RandomState(0)
    for h in (1, 3, 4, 5, 16, 17, 63, 300):
        c = rs.randint(0, 4, size=(7, h)).astype(np.uint8)
        p = lo.pack_codes(c, ld=(h + 3) // 4 + 5)
        assert p.shape == (7, (h + 3) // 4 + 5) and not p[:, (h + 3) // 4:].any()
        assert np.array_equal(lo.unpack_codes(p, h), c)
    assert lo.pack_codes(np.uint8([[1, 2, 3, 0, 2]])).tolist() == [[0b00111001, 0b00000010]]
