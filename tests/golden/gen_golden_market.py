"""
Writes tests/golden/market_*.npz by driving the UNMODIFIED reference classes
envs/market_envs.py Market_Inv{A,B,C}_{D1,Dx} (through oracle/ref_shim.py, in
the build container only) over seeded price histories and actions, feeding them
the way scripts/rl_market.py:209-239 does: reset(observed_market_state(extract,
0, ...)), then step(action, observed_market_state(extract, time_step, ...)).

    python tests/golden/gen_golden_market.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import golden_io  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    mod = ref_shim.load("envs.market_envs")
    res = ref_shim.load("tools.env_resources")
    for case in golden_io.MARKET_CASES:
        name, investor, history, n, d, tl, steps = case
        prices, actions = golden_io.market_inputs(case)
        cls = getattr(mod, f"Market_Inv{investor}_{'Dx' if history else 'D1'}")
        with ref_shim.reference_cwd():
            env = cls(n, tl, d)
        obs_days = d if history else 1
        episode = 0
        start = golden_io.market_episode_start(episode, len(prices), tl, d)
        extract, time_step = prices[start:], 0
        first = np.array(res.observed_market_state(extract, 0, 1, obs_days), dtype=np.float64)
        assert np.array_equal(first, golden_io.market_observed(extract, 0, obs_days))
        state0 = np.array(env.reset(first.copy()), dtype=np.float64)
        states, rewards, dones, risks = [], [], [], []
        for t in range(steps):
            time_step += 1
            nxt = np.array(res.observed_market_state(extract, time_step, 1, obs_days), dtype=np.float64)
            assert np.array_equal(nxt, golden_io.market_observed(extract, time_step, obs_days))
            ns, rew, done, risk = env.step(actions[t].copy(), nxt.copy())
            states.append(np.array(ns, dtype=np.float64).copy())
            rewards.append(float(rew))
            dones.append([bool(done[0]), bool(done[1])])
            risks.append(np.array(risk, dtype=np.float64).copy())
            if done[0]:
                episode += 1
                start = golden_io.market_episode_start(episode, len(prices), tl, d)
                extract, time_step = prices[start:], 0
                env.reset(np.array(res.observed_market_state(extract, 0, 1, obs_days), dtype=np.float64))
        out = os.path.join(HERE, f"market_{name}.npz")
        np.savez_compressed(out, state0=state0, states=np.array(states), rewards=np.array(rewards),
                            dones=np.array(dones), risks=np.array(risks))
        d_ = np.array(dones)
        print("wrote", out, np.array(states).shape, "episodes:", int(d_[:, 0].sum()), "learn_dones:",
              int(d_[:, 1].sum()), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
