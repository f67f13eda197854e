"""Pins oracle/market_oracle.py against trajectories of the unmodified reference market envs."""
import numpy as np
import pytest

import golden_io
from oracle import env_oracle as eo
from oracle import market_oracle as mo


def drive(case, make_env, to_np=lambda x: x):
    """Feeds an env the fixture's inputs the way scripts/rl_market.py does; yields per-step outputs."""
    name, investor, history, n, d, tl, steps = case
    prices, actions = golden_io.market_inputs(case)
    obs_days = d if history else 1
    env = make_env()
    episode = 0
    extract = prices[golden_io.market_episode_start(0, len(prices), tl, d):]
    time_step = 0
    state0 = to_np(env.reset(np.ascontiguousarray(golden_io.market_observed(extract, 0, obs_days))))
    out = []
    for t in range(steps):
        time_step += 1
        nxt = np.ascontiguousarray(golden_io.market_observed(extract, time_step, obs_days))
        ns, rew, done, risk = env.step(actions[t], nxt)
        out.append((to_np(ns), rew, done, to_np(risk)))
        if np.asarray(done).reshape(-1)[0]:
            episode += 1
            extract = prices[golden_io.market_episode_start(episode, len(prices), tl, d):]
            time_step = 0
            env.reset(np.ascontiguousarray(golden_io.market_observed(extract, 0, obs_days)))
    return state0, out


def check(case, state0, out, rtol):
    gold = golden_io.load("market_" + case[0])
    assert np.array_equal(np.asarray(state0).reshape(-1), gold["state0"])
    for t, (ns, rew, done, risk) in enumerate(out):
        assert np.array_equal(np.asarray(done, dtype=bool).reshape(-1), gold["dones"][t]), (t, done, gold["dones"][t])
        for got, want in ((np.asarray(ns).reshape(-1), gold["states"][t]),
                          (np.asarray(rew, dtype=np.float64).reshape(-1), gold["rewards"][t:t + 1]),
                          (np.asarray(risk).reshape(-1), gold["risks"][t])):
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=rtol, atol=0), (t, got, want)


@pytest.mark.parametrize("case", golden_io.MARKET_CASES, ids=lambda c: c[0])
def test_market_oracle_matches_reference(case):
    name, investor, history, n, d, tl, steps = case
    state0, out = drive(case, lambda: mo.BatchedMarket(investor, n, tl, d, history, 1))
    check(case, state0, out, rtol=1e-15)


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 15, 16, 17, 26, 64, 100, 128])
def test_row_sum_is_numpys_pairwise_sum(n):
    """The summation order restated for the kernels equals np.sum on each contiguous row, bit for bit."""
    rs = np.random.RandomState(n)
    x = rs.standard_normal((500, n)) * 10.0 ** rs.uniform(-3, 3, size=(500, n))
    want = np.array([np.sum(row) for row in x])
    assert np.array_equal(eo.np_sum_rows(x), want)
    assert np.array_equal(eo.np_sum_rows(x) / n, np.array([np.mean(row) for row in x]))
