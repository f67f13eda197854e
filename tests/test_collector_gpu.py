"""
GPU parity of the fused collector (env step -> replay append -> n-step sample) and
of the evaluation rollouts, through the C ABI, against the CPU oracles composed
the way the reference's loop composes its classes
(scripts/rl_multiplicative.py:185-273, tools/eval_episodes.py:231-315):
one `BatchedEnv(E=1)` + one `ReplayOracle` per lane.
"""
import numpy as np
import pytest
import torch

import golden_io
from oracle import env_oracle as eo
from oracle.replay_oracle import ReplayOracle
from test_env_gpu import close, make

pytestmark = pytest.mark.gpu


def _inputs(S, A, n_steps, dyn, lane_len, batch=64):
    return {"input_dims": (S,), "num_actions": A, "mini_batch_size": batch, "discount": 0.97, "multi_steps": n_steps,
            "r_abs_zero": None, "dynamics": dyn, "buffer": lane_len, "n_cumsteps": lane_len}


def _drive(family, investor, n_g, E, T, n_steps, dyn, seed=3):
    """Runs the collector and the per-lane oracles over T steps of identical actions / returns."""
    from rlmd_b200.collector import Collector
    rs = np.random.RandomState(seed)
    env = make(family, investor, n_g, n_envs=E)
    S, A = env.state_dim, env.action_dim
    col = Collector(env, T, _inputs(S, A, n_steps, dyn, T))
    lanes = [(eo.BatchedEnv(family, investor, n_g, 1), ReplayOracle(_inputs(S, A, n_steps, dyn, T))) for _ in range(E)]
    states = [l[0].reset()[0] for l in lanes]
    assert np.array_equal(col.state.cpu().numpy(), np.stack(states))
    for t in range(T):
        a = rs.uniform(-0.99, 0.99, size=(E, A))
        a[rs.random_sample(E) < 0.05] = 0.99            # saturated leverage ends episodes (learn_done)
        d = rs.standard_normal((E, env.n_gambles)) if family == "gbm" else rs.random_sample((E, env.n_gambles))
        r = golden_io.env_returns(family, d)
        obs = col.step(torch.from_numpy(a).cuda(), returns=torch.from_numpy(np.ascontiguousarray(r)).cuda())
        got_done = col.done.cpu().numpy().astype(bool)
        for e, (oe, orp) in enumerate(lanes):
            r_in = r[e:e + 1, 0] if family == "dice_sh" else r[e:e + 1]
            ns, rew, done, risk = oe.step(a[e:e + 1], r_in)
            orp.store_exp(states[e], a[e], float(rew[0]), ns[0], bool(done[0, 1]))
            assert np.array_equal(got_done[e], done[0]), (t, e)
            states[e] = oe.reset()[0] if done[0, 0] else ns[0]
        close(obs.cpu().numpy(), np.stack(states))
    return col, lanes


@pytest.mark.parametrize("family,investor,n_g,n_steps,dyn", [
    ("coin", "A", 1, 1, "M"), ("coin", "A", 1, 5, "M"), ("coin", "B", 2, 3, "A"), ("dice", "C", 1, 10, "M"),
    ("gbm", "A", 2, 4, "M"), ("dice_sh", "B", 1, 2, "A")])
def test_collector_memory_and_samples_match_the_lane_oracles(family, investor, n_g, n_steps, dyn):
    E, T = 6, 90
    col, lanes = _drive(family, investor, n_g, E, T, n_steps, dyn)
    # replay memory, lane by lane
    for e, (_, orp) in enumerate(lanes):
        sl = slice(e, None, E)                      # slot-major: lane e holds slots e, e + E, ...
        # exp/log (reward, GBM factor) differ from NumPy's in the last fp64 ulp: fp32 copies agree to 1 ulp
        assert np.allclose(col.state_memory[sl].cpu().numpy(), orp.state_memory, rtol=2e-7, atol=0)
        assert np.allclose(col.next_state_memory[sl].cpu().numpy(), orp.next_state_memory, rtol=2e-7, atol=0)
        assert np.array_equal(col.action_memory[sl].cpu().numpy(), orp.action_memory)
        assert np.allclose(col.reward_memory[sl].cpu().numpy(), orp.reward_memory, rtol=2e-7, atol=0)
        assert np.array_equal(col.terminal_memory[sl].cpu().numpy(), orp.terminal_memory)
        assert int(col.header[e, 0]) == orp.mem_idx == T
    # every slot of every lane as a sample: the n-step targets of the reference's bookkeeping
    B = col.batch_size
    slots = np.arange(E * T, dtype=np.int64)
    pad = (-len(slots)) % B
    slots_p = np.concatenate([slots, slots[:pad]])
    k = len(slots_p) // B
    # the rewards the oracle multiplies must be the device's (exp/log differ in the last ulp)
    for e, (_, orp) in enumerate(lanes):
        sl = slice(e, None, E)
        orp.reward_memory[:] = col.reward_memory[sl].cpu().numpy()
        orp.state_memory[:] = col.state_memory[sl].cpu().numpy()
        orp.next_state_memory[:] = col.next_state_memory[sl].cpu().numpy()
    got = [x.cpu().numpy() for x in col.sample(k, batch=slots_p.reshape(k, B))]
    for e, (_, orp) in enumerate(lanes):
        want = orp.sample_exp(np.arange(T))
        sel = slice(e, E * T, E)
        assert np.array_equal(got[0][sel], want[0]), (e, "states")
        assert np.array_equal(got[1][sel], want[1]), (e, "actions")
        assert np.array_equal(got[2][sel].view(np.uint32), want[2].view(np.uint32)), (e, "n-step reward")
        assert np.array_equal(got[3][sel], want[3])
        assert np.array_equal(got[4][sel], want[4])
        if n_steps > 1:
            assert np.array_equal(got[5][sel], want[5]), (e, "eff")


@pytest.mark.parametrize("n_steps,dyn", [(1, "M"), (5, "M"), (10, "A")])
def test_bulk_gather_over_lanes_equals_the_per_call_gather(n_steps, dyn):
    """>= 4096 samples per launch take the warp-cooperative gather: slot-major lanes, the same bits."""
    E, T = 6, 90
    col, _ = _drive("coin", "B", 2, E, T, n_steps, dyn)
    B = col.batch_size
    k = -(-4100 // B)
    slots = np.random.RandomState(3).randint(0, E * T, size=(k, B)).astype(np.int64)
    big = [x.clone() for x in col.sample(k, batch=slots)]
    for j in range(k):
        small = col.sample(1, batch=slots[j:j + 1])
        for g_, w_ in zip(big, small):
            assert torch.equal(g_[j * B:(j + 1) * B], w_), j


def test_single_lane_is_the_reference_stream():
    """n_envs = 1: the collector's memory equals ReplayBufferTorch fed by the env, transition by transition."""
    from rlmd_b200.replay_torch import ReplayBufferTorch
    col, lanes = _drive("coin", "A", 1, 1, 200, 5, "M", seed=11)
    orp = lanes[0][1]
    inputs = dict(_inputs(5, 1, 5, "M", 200), gpu="cuda:0")
    buf = ReplayBufferTorch(inputs)
    for i in range(200):
        buf.store_exp(orp.state_memory[i], orp.action_memory[i], float(col.reward_memory[i]), orp.next_state_memory[i],
                      bool(orp.terminal_memory[i]))
    idx = np.random.RandomState(0).permutation(200)[:64]
    a = buf.sample_exp(batch=idx)
    b = col.sample(1, batch=idx)
    for x, y in zip(a, b):
        assert torch.equal(x.reshape(-1), y.reshape(-1).to(x.dtype))


def test_drawn_batches_are_distinct_in_range_and_advance():
    from rlmd_b200.collector import Collector
    env = make("dice", "A", 1, n_envs=50, seed=4)
    col = Collector(env, 40, _inputs(5, 1, 3, "M", 40, batch=256), seed=9)
    act = torch.full((50, 1), 0.3, dtype=torch.float64, device="cuda")
    for _ in range(17):
        col.step(act)
    col.sample(4)
    idx1 = col.last_batch.cpu().numpy()
    col.sample(4)
    idx2 = col.last_batch.cpu().numpy()
    for row in np.concatenate([idx1, idx2]):
        assert len(np.unique(row)) == 256
        assert (row >= 0).all() and ((row // 50) < 17).all()              # only filled slots (slot = local * 50 + lane)
    assert not np.array_equal(idx1, idx2)
    assert int(col.counter[0]) == 17 and int(col.counter[1]) == 2


def test_graph_replay_equals_eager_steps():
    """step + sample captured in a CUDA graph: replays produce what eager calls produce."""
    from rlmd_b200.collector import Collector
    def run(graph):
        env = make("coin", "A", 1, n_envs=300, seed=21)
        col = Collector(env, 64, _inputs(5, 1, 4, "M", 64, batch=128), seed=5)
        act = torch.zeros((300, 1), dtype=torch.float64, device="cuda")
        outs = []
        step = col.capture(act, k=2) if graph else None
        n0 = col.steps
        for t in range(n0, 40):
            act.copy_(torch.full((300, 1), 0.05 + 0.02 * (t % 9), dtype=torch.float64, device="cuda"))
            if graph:
                o = step()
                outs.append((o["rewards"].clone(), o["idx"].clone(), col.state.clone()))
            else:
                col.step(act)
                s = col.sample(2)
                outs.append((s[2].clone(), col.last_batch.clone(), col.state.clone()))
        return outs, col
    # the capture warm-up consumes one step + one sample: give the eager run the same prefix
    g_outs, g_col = run(True)
    env = make("coin", "A", 1, n_envs=300, seed=21)
    from rlmd_b200.collector import Collector as Cc
    col = Cc(env, 64, _inputs(5, 1, 4, "M", 64, batch=128), seed=5)
    act = torch.zeros((300, 1), dtype=torch.float64, device="cuda")
    col.step(act)
    col.sample(2)
    e_outs = []
    for t in range(1, 40):
        act.copy_(torch.full((300, 1), 0.05 + 0.02 * (t % 9), dtype=torch.float64, device="cuda"))
        col.step(act)
        s = col.sample(2)
        e_outs.append((s[2].clone(), col.last_batch.clone(), col.state.clone()))
    assert len(g_outs) == len(e_outs) == 39
    for (gr, gi, gs), (er, ei, es) in zip(g_outs, e_outs):
        assert torch.equal(gi.reshape(-1), ei.reshape(-1)) and torch.equal(gr, er) and torch.equal(gs, es)
    assert torch.equal(g_col.header, col.header) and torch.equal(g_col.reward_memory, col.reward_memory)


# ------------------------------------------------------------------ rollouts
@pytest.mark.parametrize("family,investor,n_g", [("coin", "A", 1), ("coin", "C", 2), ("dice", "B", 1), ("gbm", "A", 3),
                                                 ("dice_sh", "I", 1), ("dice_sh", "C", 1)])
def test_rollout_matches_stepping_the_oracle(family, investor, n_g):
    from rlmd_b200 import collector
    E, T = 700, 60
    rs = np.random.RandomState(5)
    env = make(family, investor, n_g, n_envs=1)
    ref = eo.BatchedEnv(family, investor, n_g, E)
    a = rs.uniform(-0.6, 0.6, size=(E, ref.A))
    a[:20] = 0.99
    d = rs.standard_normal((T, E, ref.n)) if family == "gbm" else rs.random_sample((T, E, ref.n))
    r = np.stack([golden_io.env_returns(family, d[t]) for t in range(T)])
    reward, steps, risk, last = collector.rollout(env, a, T, returns=r)
    ref.reset()
    w_reward, w_steps = np.zeros(E), np.zeros(E, dtype=np.int64)
    w_risk, w_last = np.zeros((E, ref.R)), np.zeros((E, ref.S))
    alive = np.ones(E, dtype=bool)
    for t in range(T):
        r_in = r[t][:, 0] if family == "dice_sh" else r[t]
        ns, rew, done, rk = ref.step(a, r_in)
        w_reward[alive], w_risk[alive], w_last[alive] = rew[alive], rk[alive], ns[alive]
        w_steps[alive] += 1
        alive &= ~done[:, 0]
    assert np.array_equal(steps.cpu().numpy(), w_steps)
    close(reward.cpu().numpy(), w_reward)
    close(risk.cpu().numpy(), w_risk)
    close(last.cpu().numpy(), w_last)
    assert (w_steps < T).any() and (w_steps == T).any()           # both exits are exercised


def test_rollout_philox_equals_stepping_the_batched_env():
    from rlmd_b200 import collector, envs
    E, T = 4096, 30
    a = torch.linspace(-0.9, 0.9, E, dtype=torch.float64, device="cuda").reshape(E, 1)
    env = envs.Dice_InvA(1, n_envs=E, seed=77)
    reward, steps, risk, last = collector.rollout(env, a, T, draw_base=0)
    env.reset()
    alive = torch.ones(E, dtype=torch.bool, device="cuda")
    w_reward = torch.zeros(E, dtype=torch.float64, device="cuda")
    w_steps = torch.zeros(E, dtype=torch.int32, device="cuda")
    for t in range(T):
        ns, rew, done, rk = env.step(a)                                  # draw_index = t
        w_reward = torch.where(alive, rew, w_reward)
        w_steps += alive.int()
        alive &= ~done[:, 0]
    assert torch.equal(steps, w_steps) and torch.equal(reward, w_reward)


def test_eval_multiplicative_summary(capsys):
    from rlmd_b200 import collector, envs
    env = envs.Coin_InvA(1, seed=3)
    out = collector.eval_multiplicative(env, np.array([0.25]), n_eval=1000, max_eval_steps=200)
    assert out["reward"].shape == (1000,) and out["risk"].shape == (1000, 4)
    want = collector.eval_summary(out["reward"], out["risk"][:, 1], out["risk"][:, 3], out["steps"])
    assert np.allclose(out["stats"], want)
    # numpy's own percentile on the same episodes (the reference's call)
    assert out["stats"][3] == (np.percentile(out["reward"], q=5, method="median_unbiased") - 1) * 100
    assert abs(out["stats"][0] - 0.25 * 2.0 * 100) < 1e-9            # mean leverage = a * eta
    text = capsys.readouterr().out
    assert "Summary" in text and "mean/med/95/mad/std: g%" in text
