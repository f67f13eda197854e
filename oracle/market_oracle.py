"""
TEST INFRASTRUCTURE ONLY - NumPy restatement of the reference's market
environments (envs/market_envs.py; done flags tools/env_resources.py:140-200)
for E lock-step copies.  Nothing in rlmd_b200/ may import this module.

Parity pin: tests/test_oracle_market.py checks it against
tests/golden/market_*.npz, trajectories of the UNMODIFIED reference classes
written by tests/golden/gen_golden_market.py.

Follows: constants :37-50,:67; __init__ :92-131 (and the five siblings); step
:133-202 (A_D1), :283-358 (B_D1), :440-528 (C_D1), :611-682 / :765-841 /
:924-1013 (Dx: the returns of the whole observed history enter the state, the
first n_assets of them the portfolio return); reset :204-222, :684-703.
`self.assets` is set by reset() only, so every return is relative to the
episode's first observation - kept.
"""
import numpy as np

from oracle.env_oracle import np_sum_rows

MAX_VALUE = 1e34
INITIAL_VALUE = 1e4
MIN_VALUE = max(1e-2 * INITIAL_VALUE, 1)
MAX_VALUE_RATIO = 1
MAX_ABS_ACTION = 0.99
MIN_REWARD = 1e-3
MIN_RETURN = -0.9
MAX_RETURN = 1e10
MIN_WEIGHT = 1e-5
LEV_FACTOR = 3


def dims(investor: str, n_assets: int, obs_days: int):
    extra = {"A": 0, "B": 1, "C": 2}[investor]
    risk = (3 if n_assets == 1 else 4) + extra + n_assets
    return 4 + obs_days * n_assets, extra + n_assets, risk


class BatchedMarket:
    def __init__(self, investor: str, n_assets: int, time_length: int, obs_days: int, history: bool, n_envs: int = 1):
        self.investor, self.n, self.E = investor, int(n_assets), int(n_envs)
        self.D = int(obs_days) if history else 1
        self.time_length = int(time_length) - int(obs_days) + 1 if history else int(time_length)
        self.S, self.A, self.R = dims(investor, self.n, self.D)
        self.W = self.D * self.n
        self.wealth = np.full(self.E, INITIAL_VALUE)
        self.time = np.ones(self.E, dtype=np.int64)
        self.assets = np.ones((self.E, self.W))

    def reset(self, assets, mask=None):
        assets = np.asarray(assets, dtype=np.float64).reshape(self.E, self.W)
        m = np.ones(self.E, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        self.wealth = np.where(m, INITIAL_VALUE, self.wealth)
        self.time = np.where(m, 1, self.time)
        self.assets = np.where(m[:, None], assets, self.assets)
        state = np.zeros((self.E, self.S))
        state[:, 0:4] = [INITIAL_VALUE, 0, 1, 1]
        if self.D > 1:
            state[:, 4:] = assets
        return state / MAX_VALUE

    def step(self, action, next_assets):
        E, n = self.E, self.n
        action = np.asarray(action, dtype=np.float64).reshape(E, self.A)
        nxt = np.asarray(next_assets, dtype=np.float64).reshape(E, self.W)
        o = {"A": 0, "B": 1, "C": 2}[self.investor]
        lev = action[:, o:] * LEV_FACTOR
        hist = nxt / self.assets - 1
        r = hist[:, :n]
        step_return = np.clip(np_sum_rows(lev * r), MIN_RETURN, MAX_RETURN)
        w0 = self.wealth
        stop = retention = active = None
        if self.investor == "A":
            wmin = np.full(E, MIN_VALUE)
            wealth = np.clip(w0 * (1 + step_return), MIN_VALUE, MAX_VALUE)
        else:
            stop = (action[:, 0] + MAX_ABS_ACTION) / 2
            floor_b = np.maximum(INITIAL_VALUE * stop, MIN_VALUE)
            if self.investor == "B":
                wmin = floor_b
            else:
                retention = (action[:, 1] + MAX_ABS_ACTION) / 2
                wmin = np.where(w0 <= INITIAL_VALUE, floor_b, INITIAL_VALUE + (w0 - INITIAL_VALUE) * retention)
            active = np.maximum(w0 - wmin, 0)
            wealth = np.clip(wmin + active * (1 + step_return), wmin, MAX_VALUE)
        growth = wealth / INITIAL_VALUE
        with np.errstate(divide="ignore", invalid="ignore"):
            reward = np.exp(np.log(growth) / self.time)
        next_state = np.concatenate([np.stack([wealth, step_return, growth, reward], axis=1), hist], axis=1) / MAX_VALUE

        abs_lev = np.abs(lev)
        done_time = self.time == self.time_length
        done_state = (next_state >= MAX_VALUE_RATIO).any(axis=1)
        done = done_time | (wealth == wmin) | (reward < MIN_REWARD) | (step_return == MIN_RETURN) \
            | (abs_lev == MAX_ABS_ACTION * LEV_FACTOR).all(axis=1) | (abs_lev < MIN_WEIGHT).all(axis=1) | done_state
        if active is not None:
            done = done | (active == 0)
        learn_done = done & ~(done_time | done_state)

        cols = [reward, wealth, step_return, np_sum_rows(lev) / n]
        if stop is not None:
            cols.append(stop)
        if retention is not None:
            cols.append(retention)
        if n > 1:
            cols += [lev[:, i] for i in range(n)]
        self.wealth = wealth
        self.time = self.time + 1
        return next_state, reward, np.stack([done, learn_done], axis=1), np.stack(cols, axis=1)
