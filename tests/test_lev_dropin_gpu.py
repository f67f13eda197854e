"""
Drop-in tests: rlmd_b200.lev_exp called exactly the way tests/golden/gen_golden.py
called the reference's lev.lev_exp, compared with what the reference returned.
"""
import contextlib
import io

import numpy as np
import pytest
import torch as T

import golden_io
from oracle import lev_oracle as lo
from test_oracle_lev import assert_stats_close

pytestmark = pytest.mark.gpu


def call_smart(lx, case, outcomes):
    n, h, top, v0 = case["n"], case["h"], case["top"], case["v0"]
    dev = T.device("cuda:0")
    common = (T.tensor(n, dtype=T.int32), T.tensor(h, dtype=T.int32), top, T.tensor(v0))
    grid = case["grid"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        if case["kind"] == "coin":
            r = lx.coin_smart_lev(dev, T.tensor(outcomes.astype(np.float32)), *common, case["up_r"], case["down_r"], *grid)
        elif case["kind"] == "dice":
            r = lx.dice_smart_lev(dev, T.tensor(outcomes.astype(np.int64)), *common, case["up_r"], case["down_r"],
                                  case["mid_r"], *grid)
        elif case["kind"] == "dice_sh":
            r = lx.dice_sh_smart_lev(dev, T.tensor(outcomes.astype(np.int64)), *common, case["up_r"], case["down_r"],
                                     case["mid_r"], *case["sh"], *grid)
        else:
            r = lx.gbm_smart_lev(dev, T.tensor(outcomes), *common, *grid)
    return r, buf.getvalue()


@pytest.mark.parametrize("case", golden_io.LEV_CASES, ids=lambda c: c["name"])
def test_smart_lev_dropin(case):
    from rlmd_b200 import lev_exp
    oc = golden_io.draw_outcomes(case)
    gold = golden_io.load("lev_" + case["name"], oc)
    (data, data_T), text = call_smart(lev_exp, case, oc)
    assert data.dtype == T.float32 and data_T.dtype == T.float32
    assert tuple(data.shape) == (gold["data_T"].shape[0], 13, case["h"] - 1)
    assert tuple(data_T.shape) == gold["data_T"].shape
    data = data.cpu().numpy()[:, :, gold["cols"]]
    data_T = data_T.cpu().numpy()
    want, want_T = gold["data"], gold["data_T"]
    assert np.array_equal(data[:, 12], want[:, 12])          # leverage row
    assert text.count("lev ") == want.shape[0]
    if case["kind"] != "gbm":
        assert np.array_equal(data_T.view(np.uint32), want_T.view(np.uint32))
        assert np.array_equal(data[:, 9:12].view(np.uint32), want[:, 9:12].view(np.uint32))   # medians: bit-exact
        assert_stats_close(data[:, :9], want[:, :9])
    else:
        valid = np.isfinite(want_T) & (want_T > 1e-30)
        rel = np.abs(data_T[valid].astype(np.float64) - want_T[valid]) / want_T[valid]
        assert rel.max() <= 1e-5 * max(1.0, case["h"] / 500)
        if case["name"] != "gbm_overflow":
            assert_stats_close(data[:, :12], want[:, :12], rtol=2e-5 * max(1.0, case["h"] / 500))
        else:
            normal = np.abs(want[:, :12]) > 1e-30
            assert_stats_close(np.where(normal, data[:, :12], 0), np.where(normal, want[:, :12], 0), rtol=2e-4)


@pytest.mark.parametrize("chunk", [32, 64, 96])
def test_series_is_chunk_invariant(chunk):
    """Any chunking of the horizon gives the same data / data_T, bit for bit."""
    from rlmd_b200 import engine, lev_exp
    case = golden_io.lev_case("dice_top5")
    oc = golden_io.draw_outcomes(case)
    lev = lo.lev_grid(*case["grid"], case["up_r"], case["down_r"])
    f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
    codes = engine.encode_codes(oc)
    a = engine.lev_series("discrete", f, lev, 100.0, 5, outcomes=codes)
    b = engine.lev_series("discrete", f, lev, 100.0, 5, outcomes=codes, chunk_steps=chunk)
    assert T.equal(a[1], b[1])
    assert T.equal(a[0][:, 9:13], b[0][:, 9:13])
    assert T.allclose(a[0], b[0], rtol=1e-6, atol=0)


def test_series_philox_matches_streamed():
    from rlmd_b200 import engine, lev_exp
    n, h, probs = 3000, 130, (1 / 6, 1 / 6, 2 / 3)
    lev = lo.lev_grid(0.1, 1.0, 0.1)
    f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
    drawn = engine.lev_draw("discrete", n, h, seed=11, probs=probs)
    a = engine.lev_series("discrete", f, lev, 100.0, 2, outcomes=drawn, chunk_steps=64)
    b = engine.lev_series("discrete", f, lev, 100.0, 2, n_investors=n, horizon=h, seed=11, probs=probs, chunk_steps=64)
    assert T.equal(a[0], b[0]) and T.equal(a[1], b[1])
    x = engine.lev_draw("gbm", n, h, seed=12, log_mean=-0.05, sigma=0.447)
    levg = lo.lev_grid(-1.0, 1.0, 0.2)
    a = engine.lev_series("gbm", levg, levg, 100.0, 2, outcomes=x, chunk_steps=64)
    b = engine.lev_series("gbm", levg, levg, 100.0, 2, n_investors=n, horizon=h, seed=12, log_mean=-0.05,
                          sigma=0.447, chunk_steps=64)
    assert T.equal(a[0], b[0]) and T.equal(a[1], b[1])
    # and the series' last step equals the final-time sweep
    c = engine.lev_sweep("gbm", levg, 100.0, outcomes=x, mode="log")["data_T"]
    assert T.equal(a[1], c)


def test_fixed_final_prints_reference_text():
    """*_fixed_final_lev prints; compare with the oracle's rendering of the chain statistics."""
    from rlmd_b200 import lev_exp
    case = golden_io.lev_case("dice_top5")
    oc = golden_io.draw_outcomes(case)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        lev_exp.dice_fixed_final_lev(T.device("cuda:0"), T.tensor(oc.astype(np.int64)), case["top"], T.tensor(1e2),
                                     0.5, -0.5, 0.05, *case["grid"])
    lev = lo.lev_grid(*case["grid"], 0.5, -0.5)
    stats = lo.fixed_final_lev(oc, lo.dice_factors(lev, 0.5, -0.5, 0.05), case["top"], 1e2)
    want = lo.format_final(lev, stats.astype(np.float32))
    got = buf.getvalue().rstrip("\n")
    # identical up to the last printed digit of mad/std of degenerate (all-equal) groups
    gl, wl = got.splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    same = sum(a == b for a, b in zip(gl, wl))
    assert same >= len(gl) - 2, (got, want)
