"""
TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy fp64, vectorised over a batch
of E independent environments) of the reference's multiplicative gym
environments.  The product (rlmd_b200/) never imports it.

Restated (reference file:line):
  constants                envs/coin_flip_envs.py:40-93, envs/dice_roll_envs.py:39-96,
                           envs/gbm_envs.py:43-90, envs/dice_roll_sh_envs.py:39-118
  Inv A / B / C step+reset envs/coin_flip_envs.py:150-233, 290-379, 436-538
  dice sampling / stop map envs/dice_roll_envs.py:174-176, 311
  GBM step                 envs/gbm_envs.py:147-212, 304-323, 449-474
  safe-haven steps         envs/dice_roll_sh_envs.py:160-235, 290-365, 420-502, 557-645
  done flags               tools/env_resources.py:26-137 (multi_dones / multi_gbm_dones)

Parity pin: tests/golden/env_*.npz hold trajectories of the UNMODIFIED reference
classes driven with injected random draws (tests/golden/gen_golden_more.py);
tests/test_oracle_env.py checks this file against them to 1e-15 relative (exact
for everything but the exp/log in the reward).  `Dice_SH_INSURED.step` builds its
`risk` vector from scalars mixed with 1-element arrays (:178-179, :228-231), which
numpy 2.x refuses; its fixture (env_dicesh_I.npz) comes from the same unmodified class
with the module's `np.array` reading such elements as scalars, as the reference's own
numpy 1.22 pin does (tests/golden/gen_golden_more.py:_Numpy122) - so it is pinned too.

Returns are INJECTED (`r`): the reference draws from the unseeded global
np.random state, which is not reproduced.
"""
from __future__ import annotations

import numpy as np

MAX_VALUE = 1e18
INITIAL_VALUE = 1e4
MIN_VALUE = max(1e-2 * INITIAL_VALUE, 1)
MAX_ABS_ACTION = 0.99
MAX_RETURN = 1e10
MIN_WEIGHT = 1e-5
MAX_VALUE_RATIO = 1

FAMILIES = {
    # name: dict(returns, probs, eta, min_reward, min_return)
    "coin": dict(returns=(0.5, -0.4), probs=(0.5, 0.5), eta=1 / 0.5, min_reward=1e-3, min_return=-0.9),
    "dice": dict(returns=(0.5, -0.5, 0.05), probs=(1 / 6, 1 / 6, 1 - (1 / 6 + 1 / 6)), eta=1 / 0.5,
                 min_reward=1e-3, min_return=-0.9),
    "gbm": dict(drift=0.0540025395205692, vol=0.1897916175617430, eta=5.0, min_reward=1e-3,
                min_return=float(np.log(0.1))),
    "dice_sh": dict(returns=(0.5, -0.5, 0.05), probs=(1 / 6, 1 / 6, 1 - (1 / 6 + 1 / 6)), eta=1 / 0.5,
                    sh=(max(-1, -0.99), 5.0, max(-1, -0.99)), i_eta=(-1 - 5) / (-0.5 - 5), sh_eta=1.0,
                    min_reward=1e-6, min_return=-0.99),
}
FAMILIES["gbm"]["log_mean"] = FAMILIES["gbm"]["drift"] - FAMILIES["gbm"]["vol"] ** 2 / 2


def dims(family: str, investor: str, n_gambles: int):
    """(state dim, action dim, risk dim)."""
    if family == "dice_sh":
        return 6, {"I": 1, "A": 2, "B": 3, "C": 4}[investor], 7
    extra = {"A": 0, "B": 1, "C": 2}[investor]
    risk = 3 + extra + n_gambles if n_gambles == 1 else 4 + extra + n_gambles
    return 4 + n_gambles, extra + n_gambles, risk


def np_sum_rows(x: np.ndarray) -> np.ndarray:
    """
    Row sums of x [E,n] in the order np.sum uses on ONE contiguous fp64 row (the
    reference sums per-env vectors): NumPy's pairwise_sum - left to right below
    8 elements, 8 running partial sums from 8 on (n <= 128).  Checked against
    np.sum itself in tests/test_oracle_env.py.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[1]
    if n < 8:
        res = x[:, 0].copy()
        for i in range(1, n):
            res = res + x[:, i]
        return res
    r = [x[:, j].copy() for j in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] = r[j] + x[:, i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res = res + x[:, i]
        i += 1
    return res


class BatchedEnv:
    """E lock-step copies of one reference env class (family, investor, n_gambles)."""

    def __init__(self, family: str, investor: str, n_gambles: int = 1, n_envs: int = 1):
        self.family, self.investor = family, investor
        self.n = 1 if family == "dice_sh" else int(n_gambles)
        self.E = int(n_envs)
        self.c = FAMILIES[family]
        self.S, self.A, self.R = dims(family, investor, self.n)
        self.reset()

    def reset(self, mask=None):
        if mask is None:
            self.wealth = np.full(self.E, INITIAL_VALUE)
            self.time = np.ones(self.E, dtype=np.int64)
        else:
            self.wealth = np.where(mask, INITIAL_VALUE, self.wealth)
            self.time = np.where(mask, 1, self.time)
        state = np.zeros((self.E, self.S))
        state[:, 0:4] = [INITIAL_VALUE, 0, 1, 1]
        return state / MAX_VALUE

    def step(self, action: np.ndarray, r: np.ndarray):
        """
        action [E,A] fp64; r [E,n] injected returns (dice_sh: [E] the die return;
        the safe-haven return follows from it).  Returns next_state [E,S],
        reward [E], done [E,2] bool, risk [E,R].
        """
        c = self.c
        E, n = self.E, self.n
        action = np.asarray(action, dtype=np.float64).reshape(E, self.A)
        w0 = self.wealth
        inv = self.investor
        nan = np.full(E, np.nan)
        stop = retention = None

        if self.family == "dice_sh":
            r = np.asarray(r, dtype=np.float64).reshape(E)
            up, down, mid = c["returns"]
            r_sh = np.where(r == mid, c["sh"][2], np.where(r == up, c["sh"][0], c["sh"][1]))
            if inv == "I":
                lev = action[:, 0] * c["i_eta"]
                lev_sh = 1 - lev
            else:
                o = {"A": 0, "B": 1, "C": 2}[inv]
                lev = action[:, o] * c["eta"]
                lev_sh = (action[:, o + 1] + MAX_ABS_ACTION) / 2 * c["sh_eta"]
            step_return = np.clip(lev * r + lev_sh * r_sh, c["min_return"], MAX_RETURN)
            factor = 1 + step_return
            levs = lev[:, None]
            rs = np.stack([r, r_sh], axis=1)
        else:
            o = {"A": 0, "B": 1, "C": 2}[inv]
            r = np.asarray(r, dtype=np.float64).reshape(E, n)
            levs = action[:, o:] * c["eta"]
            total = np_sum_rows(levs * r)
            if self.family == "gbm":
                step_return = np.maximum(total, c["min_return"])
                factor = np.minimum(np.exp(step_return), 1 + MAX_RETURN)
            else:
                step_return = np.clip(total, c["min_return"], MAX_RETURN)
                factor = 1 + step_return
            rs = r

        if inv in ("A", "I"):
            wmin = np.full(E, MIN_VALUE)
            active = None
            wealth = np.clip(w0 * factor, MIN_VALUE, MAX_VALUE)
        else:
            if inv == "B" and self.family == "coin":
                stop = np.abs(action[:, 0])                      # envs/coin_flip_envs.py:308
            else:
                stop = (action[:, 0] + MAX_ABS_ACTION) / 2
            floor_b = np.maximum(INITIAL_VALUE * stop, MIN_VALUE)
            if inv == "B":
                wmin = floor_b
            else:
                retention = (action[:, 1] + MAX_ABS_ACTION) / 2
                wmin = np.where(w0 <= INITIAL_VALUE, floor_b, INITIAL_VALUE + (w0 - INITIAL_VALUE) * retention)
            active = np.maximum(w0 - wmin, 0)
            wealth = np.clip(wmin + active * factor, wmin, MAX_VALUE)

        growth = wealth / INITIAL_VALUE
        with np.errstate(divide="ignore", invalid="ignore"):
            reward = np.exp(np.log(growth) / self.time)

        next_state = np.concatenate([np.stack([wealth, step_return, growth, reward], axis=1), rs], axis=1)
        next_state = next_state / MAX_VALUE

        # tools/env_resources.py:26-137
        abs_lev = np.abs(levs)
        hit = abs_lev == MAX_ABS_ACTION * c["eta"]
        lev_max = hit.all(axis=1) if self.family == "gbm" else hit.any(axis=1)
        lev_min = (abs_lev < MIN_WEIGHT).all(axis=1)
        done_state = (next_state >= MAX_VALUE_RATIO).any(axis=1)
        done = (wealth == wmin) | (reward < c["min_reward"]) | (step_return == c["min_return"]) | lev_max \
            | lev_min | done_state
        if active is not None:
            done = done | (active == 0)
        learn_done = done & ~done_state

        mean_lev = np_sum_rows(levs) / levs.shape[1]
        if self.family == "dice_sh":
            cols = [reward, wealth, step_return, levs[:, 0], stop if stop is not None else nan,
                    retention if retention is not None else nan, lev_sh]
        else:
            cols = [reward, wealth, step_return, mean_lev]
            if stop is not None:
                cols.append(stop)
            if retention is not None:
                cols.append(retention)
            if n > 1:
                cols += [levs[:, i] for i in range(n)]
        risk = np.stack(cols, axis=1)

        self.wealth = wealth
        self.time = self.time + 1
        return next_state, reward, np.stack([done, learn_done], axis=1), risk
