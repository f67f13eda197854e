"""Pins oracle/env_oracle.py against trajectories of the unmodified reference env classes."""
import numpy as np
import pytest

import golden_io
from oracle import env_oracle as eo


def replay_case(make_env, gold, family):
    """Drive `make_env()` through the fixture's actions/returns; yields per-step outputs."""
    env = make_env()
    out = []
    for t in range(golden_io.ENV_STEPS):
        r = gold["returns"][t]
        ns, rew, done, risk = env.step(gold["actions"][t][None, :], r[None, :] if family != "dice_sh" else r)
        out.append((ns[0], rew[0], done[0], risk[0]))
        if done[0, 0]:
            env.reset()
    return out


@pytest.mark.parametrize("case", golden_io.ENV_CASES, ids=lambda c: c[0])
def test_env_oracle_matches_reference(case):
    name, _, _, family, investor, n_g = case
    gold = golden_io.load("env_" + name)
    env0 = eo.BatchedEnv(family, investor, n_g, 1)
    assert np.array_equal(env0.reset()[0], gold["state0"])
    out = replay_case(lambda: eo.BatchedEnv(family, investor, n_g, 1), gold, family)
    for t, (ns, rew, done, risk) in enumerate(out):
        assert np.array_equal(done, gold["dones"][t]), (t, done, gold["dones"][t])
        for got, want in ((ns, gold["states"][t]), (np.array([rew]), gold["rewards"][t:t + 1]), (risk, gold["risks"][t])):
            assert got.shape == want.shape
            assert np.array_equal(np.isnan(got), np.isnan(want))
            ok = ~np.isnan(want)
            assert np.allclose(got[ok], want[ok], rtol=1e-15, atol=0), (t, got, want)
    assert gold["resets"].sum() >= 20   # the fixtures do exercise terminations


def test_env_oracle_batch_equals_singles():
    """The vectorised oracle steps E envs exactly like E single envs."""
    rs = np.random.RandomState(0)
    E, T = 64, 30
    for family, investor, n_g in (("coin", "C", 2), ("gbm", "B", 3), ("dice_sh", "C", 1), ("dice_sh", "I", 1)):
        big = eo.BatchedEnv(family, investor, n_g, E)
        singles = [eo.BatchedEnv(family, investor, n_g, 1) for _ in range(E)]
        for t in range(T):
            a = rs.uniform(-0.99, 0.99, size=(E, big.A))
            d = rs.standard_normal((E, big.n)) if family == "gbm" else rs.random_sample((E, big.n))
            r = golden_io.env_returns(family, d)
            r_in = r[:, 0] if family == "dice_sh" else r
            ns, rew, done, risk = big.step(a, r_in)
            for e in range(E):
                s1 = singles[e].step(a[e:e + 1], r_in[e:e + 1])
                assert np.array_equal(s1[0][0], ns[e]) and s1[1][0] == rew[e] and np.array_equal(s1[2][0], done[e])
                assert np.array_equal(s1[3][0], risk[e], equal_nan=True)
            big.reset(done[:, 0])
            for e in range(E):
                if done[e, 0]:
                    singles[e].reset()
