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
