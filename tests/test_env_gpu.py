"""
GPU parity of the batched multiplicative env step (K4) through the C ABI:
(a) the reference's own trajectories (tests/golden/env_*.npz), env by env;
(b) large batches against the vectorised CPU oracle on the same injected returns.
"""
import numpy as np
import pytest
import torch

import golden_io
from oracle import env_oracle as eo
from oracle import philox_oracle as po

pytestmark = pytest.mark.gpu


def make(family, investor, n_g, **kw):
    from rlmd_b200 import envs
    name = {"coin": "Coin_Inv", "dice": "Dice_Inv", "gbm": "GBM_Inv"}.get(family)
    if family == "dice_sh":
        cls = getattr(envs, "Dice_SH_INSURED" if investor == "I" else f"Dice_SH_Inv{investor}")
        return cls(**kw)
    return getattr(envs, name + investor)(n_g, **kw)


def close(got, want, rtol=1e-12):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.allclose(got[ok], want[ok], rtol=rtol, atol=0), np.abs(got[ok] - want[ok]).max()


@pytest.mark.parametrize("case", golden_io.ENV_CASES, ids=lambda c: c[0])
def test_single_env_follows_reference_trajectory(case):
    name, _, _, family, investor, n_g = case
    gold = golden_io.load("env_" + name)
    env = make(family, investor, n_g)
    assert np.array_equal(env.reset(), gold["state0"])
    assert env.observation_space.shape == gold["state0"].shape
    assert env.action_space.shape == gold["actions"][0].shape
    for t in range(golden_io.ENV_STEPS):
        ns, rew, done, risk = env.step(gold["actions"][t], returns=gold["returns"][t])
        assert done == list(gold["dones"][t]), (t, done)
        close(ns, gold["states"][t])
        close([rew], gold["rewards"][t:t + 1])
        close(risk, gold["risks"][t])
        if done[0]:
            env.reset()


@pytest.mark.parametrize("family,investor,n_g", [
    ("coin", "A", 1), ("coin", "B", 3), ("coin", "C", 8), ("dice", "A", 2), ("dice", "C", 1), ("gbm", "A", 1),
    ("gbm", "B", 4), ("gbm", "C", 2), ("dice_sh", "I", 1), ("dice_sh", "A", 1), ("dice_sh", "B", 1), ("dice_sh", "C", 1)])
def test_batch_matches_oracle(family, investor, n_g):
    E, T = 5003, 25
    rs = np.random.RandomState(7)
    env = make(family, investor, n_g, n_envs=E)
    ref = eo.BatchedEnv(family, investor, n_g, E)
    for t in range(T):
        a = rs.uniform(-0.99, 0.99, size=(E, ref.A))
        a[rs.random_sample(E) < 0.02] = 0.99          # saturated leverage
        a[rs.random_sample(E) < 0.02] = 1e-8          # vanishing leverage
        d = rs.standard_normal((E, ref.n)) if family == "gbm" else rs.random_sample((E, ref.n))
        r = golden_io.env_returns(family, d)
        r_in = r[:, 0] if family == "dice_sh" else r
        ns, rew, done, risk = env.step(torch.from_numpy(a).cuda(), returns=torch.from_numpy(r).cuda())
        wns, wrew, wdone, wrisk = ref.step(a, r_in)
        assert np.array_equal(done.cpu().numpy(), wdone)                     # exact flags
        close(ns.cpu().numpy(), wns)
        close(rew.cpu().numpy(), wrew)
        close(risk.cpu().numpy(), wrisk)
        mask = wdone[:, 0]
        st = env.reset(torch.from_numpy(mask).cuda())
        wst = ref.reset(mask)
        assert np.array_equal(st.cpu().numpy()[mask], wst[mask])
        assert np.array_equal(env.time.cpu().numpy(), ref.time)
        close(env.wealth.cpu().numpy(), ref.wealth)


def test_philox_draws_are_reproducible_and_distributed():
    from rlmd_b200 import envs
    E = 200_000
    a = torch.full((E, 1), 0.25, dtype=torch.float64, device="cuda")
    e1, e2 = envs.Dice_InvA(1, n_envs=E, seed=5), envs.Dice_InvA(1, n_envs=E, seed=5)
    s1, s2 = e1.step(a)[0], e2.step(a)[0]
    assert torch.equal(s1, s2)
    r = (s1[:, 4] * 1e18).cpu().numpy()
    frac = [(np.abs(r - v) < 1e-9).mean() for v in (0.5, -0.5, 0.05)]
    assert np.allclose(frac, [1 / 6, 1 / 6, 2 / 3], atol=4e-3)
    # the discrete draw is word 0 of Philox block (env, draw_index): reproduce on the CPU
    w = po.philox4x32_10(np.arange(1000, dtype=np.uint64), 0, 0, po.TAG_ENV, 5, 0)[0].astype(np.uint64)
    thr0, thr1 = int((1 / 6) * 2 ** 32), int((1 / 6 + 1 / 6) * 2 ** 32)
    want = np.where(w < thr0, 0.5, np.where(w < thr1, -0.5, 0.05))
    assert np.array_equal(r[:1000].round(12), want)
    assert not torch.equal(e1.step(a)[0][:, 4], s1[:, 4])     # the next call draws afresh
    g = envs.GBM_InvA(2, n_envs=E, seed=9)
    z = (g.step(torch.full((E, 2), 0.1, dtype=torch.float64, device="cuda"))[0][:, 4:] * 1e18).cpu().numpy()
    mu, vol = 0.0540025395205692 - 0.1897916175617430 ** 2 / 2, 0.1897916175617430
    assert abs(z.mean() - mu) < 2e-3 and abs(z.std() - vol) < 2e-3
