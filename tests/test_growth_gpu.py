"""
Growth-rate summaries (b200_growth_summary) against NumPy on the same log-wealth:
exact order statistics (min / max / quantile neighbours), 1e-12 on the fp64 means.
"""
import numpy as np
import pytest
import torch

from oracle import lev_oracle as lo

pytestmark = pytest.mark.gpu


def run(lw, dT, v0, h, q):
    from rlmd_b200 import engine
    t = torch.as_tensor(lw, device="cuda")
    d = None if dT is None else torch.as_tensor(dT, device="cuda")
    return engine.growth_summary(t, h, v0, data_T=d, quantiles=q).cpu().numpy()


def check(got, want):
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want)), (got, want)
    ok = ~np.isnan(want)
    assert np.array_equal(got[:, 0], want[:, 0])                        # valid-run counts: exact
    assert np.array_equal(got[:, 4:6][ok[:, 4:6]], want[:, 4:6][ok[:, 4:6]])   # min / max: exact
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000, 100_003])
@pytest.mark.parametrize("q", [(0.05, 0.5), (0.0, 1.0, 0.95), ()])
def test_against_numpy(n, q):
    rs = np.random.RandomState(n)
    lw = np.log(100.0) + rs.standard_normal((5, n)) * np.array([[0.1], [3.0], [40.0], [200.0], [1e-9]])
    lw[1, ::3] = lw[1, 0]                              # ties
    dT = np.exp(lw).astype(np.float32)                 # over/underflows in fp32 for the wide rows
    check(run(lw, dT, 100.0, 250, q), lo.growth_summary(lw, dT, 100.0, 250, q))
    check(run(lw, None, 100.0, 250, q), lo.growth_summary(lw, None, 100.0, 250, q))


def test_ruined_runs():
    """Factors of zero (dice at full leverage) give log wealth -inf: counted as invalid, quantiles follow numpy."""
    rs = np.random.RandomState(1)
    lw = rs.standard_normal((2, 5000))
    lw[0, rs.rand(5000) < 0.02] = -np.inf              # 2 % ruined: the 5th percentile stays finite
    lw[1, rs.rand(5000) < 0.5] = -np.inf
    got = run(lw, None, 1.0, 100, (0.05, 0.5))
    want = lo.growth_summary(lw, None, 1.0, 100, (0.05, 0.5))
    assert np.array_equal(got[:, 0], want[:, 0])
    assert np.all(np.isneginf(got[:, 1])) and np.all(np.isneginf(got[:, 4]))
    np.testing.assert_allclose(got[:, 3], want[:, 3], rtol=1e-12)
    np.testing.assert_allclose(got[0, 6:], want[0, 6:], rtol=1e-12)
    assert np.isneginf(got[1, 6])                      # numpy yields nan for (-inf, -inf) neighbours; the engine -inf
    assert np.array_equal(got[:, 5], want[:, 5])


def test_sweep_to_growth_pipeline():
    """LOG sweep -> growth summary on device, against the oracle's log wealth on the same outcomes."""
    from rlmd_b200 import engine, lev_exp
    rs = np.random.RandomState(0)
    n, h = 20_000, 1000
    oc = rs.choice(3, size=(n, h), p=[1 / 6, 1 / 6, 2 / 3]).astype(np.uint8)
    lev = np.asarray(lev_exp.param_range(0.1, 1.0, 0.1), dtype=np.float32)
    f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
    res = engine.lev_sweep("discrete", f, 100.0, outcomes=engine.encode_codes(oc), mode="log", want_log_w=True)
    got = engine.growth_summary(res["log_w"], h, 100.0, data_T=res["data_T"]).cpu().numpy()
    lw = lo.log_wealth_discrete(oc, f, 100.0)
    with np.errstate(over="ignore"):
        want = lo.growth_summary(lw, np.exp(lw).astype(np.float32), 100.0, h)
    assert np.array_equal(got[:, 0], want[:, 0])
    np.testing.assert_allclose(got[:, 1:], want[:, 1:], rtol=1e-10, atol=1e-14)


def test_gbm_state_summary_equals_row_summary():
    """gbm_growth_summary (one selection over the leverage-independent state S) against growth_summary on the
    [G,N] log-wealth rows of the same sweep: valid-run counts exact, everything else to fp64 rounding."""
    import torch
    from rlmd_b200 import engine, lev_exp
    n, h = 200_003, 700
    for grid, mu, sig in (((-1.0, 1.0, 0.2), 0.05, 0.2 ** 0.5), ((0.2, 2.0, 0.2), 0.0540025395205692, 0.1897916175617430)):
        lev = np.asarray(lev_exp.param_range(*grid), dtype=np.float32)
        kw = dict(n_investors=n, horizon=h, seed=3, log_mean=mu - sig * sig / 2, sigma=sig, mode="log")
        a = engine.lev_sweep("gbm", lev, 100.0, want_log_w=True, **kw)
        b = engine.lev_sweep("gbm", lev, 100.0, want_state=True, **kw)
        assert torch.equal(a["data_T"], b["data_T"]) and b["log_w"] is None
        want = engine.growth_summary(a["log_w"], h, 100.0, data_T=a["data_T"], quantiles=(0.05, 0.5)).cpu().numpy()
        got = engine.gbm_growth_summary(b["state"], lev, h, 100.0, quantiles=(0.05, 0.5)).cpu().numpy()
        got_T = engine.gbm_growth_summary(b["state"], lev, h, 100.0, data_T=b["data_T"], quantiles=(0.05, 0.5))
        assert np.array_equal(got_T.cpu().numpy()[:, 0], got[:, 0])          # valid runs: from data_T or re-formed
        np.testing.assert_allclose(got_T.cpu().numpy(), got, rtol=1e-13, equal_nan=True)
        assert np.array_equal(got[:, 0], want[:, 0])              # valid runs per leverage
        assert 0 < want[:, 0].min() or want[:, 0].max() > 0
        ok = ~np.isnan(want)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_allclose(got[ok], want[ok], rtol=2e-12, atol=1e-16)
        # log wealth rebuilt from the state is the sweep's log wealth
        lw = np.log(100.0) + lev.astype(np.float64)[:, None] * b["state"][0].cpu().numpy()[None, :]
        np.testing.assert_allclose(lw, a["log_w"].cpu().numpy(), rtol=1e-15, atol=1e-13)


@pytest.mark.parametrize("depth", [1, 2])
def test_gbm_philox_pipeline_equals_direct_calls(depth):
    """FinalSweepPipeline.submit_philox (sweep on one stream, statistics + growth summaries on another) returns
    what the direct calls return, step after step, at either depth."""
    import torch
    from rlmd_b200 import engine, lev_exp
    n, h, top = 50_021, 300, 5
    lev = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), dtype=np.float32)
    pipe = engine.FinalSweepPipeline("gbm", lev, 100.0, top, depth=depth)
    assert pipe.depth == depth
    got = [pipe.submit_philox(n, h, seed=40 + i, log_mean=-0.05, sigma=0.447) for i in range(4)]
    pipe.synchronize()
    for i, (st, gr) in enumerate(got):
        res = engine.lev_sweep("gbm", lev, 100.0, n_investors=n, horizon=h, seed=40 + i, log_mean=-0.05, sigma=0.447,
                               mode="log", want_state=True)
        want_st = engine.rowstats(res["data_T"], top)
        want_gr = engine.gbm_growth_summary(res["state"], lev, h, 100.0, data_T=res["data_T"])
        assert torch.equal(st[:, 9:12], want_st[:, 9:12]) and torch.allclose(st, want_st, rtol=1e-12, atol=0, equal_nan=True)
        assert torch.equal(gr[:, 0], want_gr[:, 0]) and torch.allclose(gr, want_gr, rtol=1e-12, atol=0, equal_nan=True)
    assert engine.FinalSweepPipeline("gbm", lev, 100.0, top).depth == 2       # the GBM default
    assert engine.FinalSweepPipeline("discrete", np.ones((2, 2), np.float32), 100.0, top).depth == 1
