"""
GPU parity of the market envs through the C ABI:
(a) the reference's own trajectories (tests/golden/market_*.npz), env by env;
(b) batches against the CPU oracle.
"""
import numpy as np
import pytest
import torch

import golden_io
from oracle import market_oracle as mo
from test_env_gpu import close
from test_oracle_market import check, drive

pytestmark = pytest.mark.gpu


def make(investor, history, *args, **kw):
    from rlmd_b200 import market_envs
    return getattr(market_envs, f"Market_Inv{investor}_{'Dx' if history else 'D1'}")(*args, **kw)


@pytest.mark.parametrize("case", golden_io.MARKET_CASES, ids=lambda c: c[0])
def test_single_env_follows_reference_trajectory(case):
    name, investor, history, n, d, tl, steps = case
    env = make(investor, history, n, tl, d)
    gold = golden_io.load("market_" + name)
    assert env.observation_space.shape == gold["state0"].shape
    assert env.action_space.shape == (env.action_dim,)
    state0, out = drive(case, lambda: env)
    check(case, state0, out, rtol=1e-12)          # exp/log differ from NumPy's in the last ulp


@pytest.mark.parametrize("investor,history,n,d", [("A", False, 1, 1), ("B", False, 8, 1), ("C", False, 31, 1),
                                                  ("A", True, 2, 6), ("C", True, 17, 3), ("B", True, 128, 2)])
def test_batch_matches_oracle(investor, history, n, d):
    E, T, tl = 1501, 30, 12
    rs = np.random.RandomState(n + d)
    env = make(investor, history, n, tl, d, n_envs=E)
    ref = mo.BatchedMarket(investor, n, tl, d, history, E)
    W = ref.W
    base = 50 + 100 * rs.random_sample((E, W))
    st = env.reset(torch.from_numpy(base).cuda())
    assert np.array_equal(st.cpu().numpy(), ref.reset(base))
    for t in range(T):
        a = rs.uniform(-0.99, 0.99, size=(E, ref.A))
        a[rs.random_sample(E) < 0.03] = 0.99
        a[rs.random_sample(E) < 0.03] = 1e-8
        nxt = base * np.exp(0.05 * rs.standard_normal((E, W)))
        ns, rew, done, risk = env.step(torch.from_numpy(a).cuda(), torch.from_numpy(nxt).cuda())
        wns, wrew, wdone, wrisk = ref.step(a, nxt)
        assert np.array_equal(done.cpu().numpy(), wdone), t
        close(ns.cpu().numpy(), wns)
        close(rew.cpu().numpy(), wrew)
        close(risk.cpu().numpy(), wrisk)
        mask = wdone[:, 0]
        base = np.where(mask[:, None], 50 + 100 * rs.random_sample((E, W)), base)
        st = env.reset(torch.from_numpy(base).cuda(), mask=torch.from_numpy(mask).cuda())
        wst = ref.reset(base, mask)
        assert np.array_equal(st.cpu().numpy()[mask], wst[mask])
        assert np.array_equal(env.time.cpu().numpy(), ref.time)
        close(env.wealth.cpu().numpy(), ref.wealth)
    assert wdone[:, 0].any()


def test_observed_market_state_is_the_references_slicing():
    from rlmd_b200.market_envs import observed_market_state
    rs = np.random.RandomState(0)
    extract = rs.random_sample((60, 4))
    for obs_days in (1, 3, 7):
        for step in (0, 1, 5, 20):
            want = golden_io.market_observed(extract, step, obs_days)
            assert np.array_equal(observed_market_state(extract, step, 1, obs_days), want)
            got = observed_market_state(torch.from_numpy(extract).cuda(), step, 1, obs_days)
            assert np.array_equal(got.cpu().numpy(), want)


def test_limits():
    from rlmd_b200._lib import B200Error
    with pytest.raises(B200Error):
        make("A", False, 129, 10, 1)
