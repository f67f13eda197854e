"""
Batched variant of the reference's market environments (envs/market_envs.py)
under the reference's class names and signatures:

    Market_Inv{A,B,C}_D1(n_assets, time_length, obs_days)      obs_days unused (:99)
    Market_Inv{A,B,C}_Dx(n_assets, time_length, obs_days)      state holds obs_days of prices
    reset(assets) -> state                                     (:204-222, :684-703)
    step(action, next_assets) -> (next_state, reward, [done, learn_done], risk)

plus `observed_market_state` (tools/env_resources.py:203-226), the slicing the
training loop uses to build `assets` / `next_assets` from a market extract
(scripts/rl_market.py:209-239).  `n_envs = 1` takes and returns NumPy like the
reference; `n_envs > 1` takes CUDA float64 tensors [n_envs, ...] (one market
window per environment).  Outputs are fresh arrays (the reference returns its
aliased `self.next_state` / `self.risk`).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import MarketDesc, check, lib, ptr, require_cuda, stream_ptr
from .envs import Box

# envs/market_envs.py:37-50
MAX_VALUE = 1e34
INITIAL_VALUE = 1e4
MIN_VALUE_RATIO = 1e-2
MAX_ABS_ACTION = 0.99
MIN_REWARD = 1e-3
MIN_RETURN = -0.9
MAX_RETURN = 1e10
MIN_WEIGHT = 1e-5
LEV_FACTOR = 3
MIN_VALUE = max(MIN_VALUE_RATIO * INITIAL_VALUE, 1)

_INVESTOR = {"A": _lib.INV_A, "B": _lib.INV_B, "C": _lib.INV_C}


def observed_market_state(market_extract, time_step: int, action_days: int, obs_days: int):
    """tools/env_resources.py:203-226 (works on NumPy arrays and torch tensors alike)."""
    if obs_days == 1:
        return market_extract[time_step * action_days]
    lo = time_step * action_days
    hi = lo + obs_days if time_step > 0 else obs_days
    window = market_extract[lo:hi].reshape(-1)
    return window.flip(0) if isinstance(window, torch.Tensor) else window[::-1]


class BatchedMarketEnv:
    def __init__(self, investor: str, n_assets: int, time_length: int, obs_days: int, history: bool,
                 n_envs: int = 1, device="cuda"):
        require_cuda()
        self.investor, self.n_assets, self.n_envs = investor, int(n_assets), int(n_envs)
        self.obs_days = int(obs_days) if history else 1
        self.time_length = int(time_length) - int(obs_days) + 1 if history else int(time_length)
        self.device = torch.device(device)
        d = MarketDesc()
        d.investor, d.n_assets, d.obs_days, d.time_length = (_INVESTOR[investor], self.n_assets, self.obs_days,
                                                             self.time_length)
        d.max_value, d.initial_value, d.min_value = MAX_VALUE, INITIAL_VALUE, MIN_VALUE
        d.max_abs_action, d.min_reward, d.min_return = MAX_ABS_ACTION, MIN_REWARD, MIN_RETURN
        d.max_return, d.min_weight, d.lev_factor = MAX_RETURN, MIN_WEIGHT, LEV_FACTOR
        self._d = d
        s, a, r = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.b200_market_dims(C.byref(d), C.byref(s), C.byref(a), C.byref(r)))
        self.state_dim, self.action_dim, self.risk_dim = s.value, a.value, r.value
        self.width = self.obs_days * self.n_assets
        self.reward_range = (MIN_REWARD, np.inf)
        self.observation_space = Box(-np.inf, np.inf, (self.state_dim,))
        self.action_space = Box(-MAX_ABS_ACTION, MAX_ABS_ACTION, (self.action_dim,))
        e = self.n_envs
        with torch.cuda.device(self.device):
            self._wealth = torch.full((e,), INITIAL_VALUE, dtype=torch.float64, device=self.device)
            self._time = torch.ones(e, dtype=torch.int32, device=self.device)
            self._assets = torch.ones((e, self.width), dtype=torch.float64, device=self.device)

    @property
    def wealth(self):
        return float(self._wealth[0]) if self.n_envs == 1 else self._wealth

    @property
    def time(self):
        return int(self._time[0]) if self.n_envs == 1 else self._time

    def _dev(self, x, cols):
        t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64) if not isinstance(x, torch.Tensor) else x)
        return t.to(device=self.device, dtype=torch.float64).reshape(self.n_envs, cols).contiguous()

    def reset(self, assets, mask: Optional[torch.Tensor] = None):
        e = self.n_envs
        with torch.cuda.device(self.device):
            a = self._dev(assets, self.width)
            state = torch.empty((e, self.state_dim), dtype=torch.float64, device=self.device)
            m = None
            if mask is not None:
                m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
                state.fill_(float("nan"))
            check(lib.b200_market_reset(C.byref(self._d), e, ptr(self._wealth), ptr(self._time), ptr(a),
                                        ptr(self._assets), ptr(state), ptr(m), stream_ptr()))
        return state[0].cpu().numpy() if e == 1 else state

    def step(self, action, next_assets):
        e, dev = self.n_envs, self.device
        with torch.cuda.device(dev):
            a = self._dev(action, self.action_dim)
            nxt = self._dev(next_assets, self.width)
            ns = torch.empty((e, self.state_dim), dtype=torch.float64, device=dev)
            rew = torch.empty(e, dtype=torch.float64, device=dev)
            done = torch.empty((e, 2), dtype=torch.uint8, device=dev)
            risk = torch.empty((e, self.risk_dim), dtype=torch.float64, device=dev)
            check(lib.b200_market_step(C.byref(self._d), e, ptr(self._wealth), ptr(self._time), ptr(self._assets),
                                       ptr(a), ptr(nxt), ptr(ns), ptr(rew), ptr(done), ptr(risk), stream_ptr()))
        if e == 1:
            d = done.cpu().numpy()[0]
            return ns[0].cpu().numpy(), float(rew[0]), [bool(d[0]), bool(d[1])], risk[0].cpu().numpy()
        return ns, rew, done.bool(), risk


def _named(name, investor, history):
    class _Env(BatchedMarketEnv):
        def __init__(self, n_assets: int, time_length: int, obs_days: int, n_envs: int = 1, device="cuda"):
            super().__init__(investor, n_assets, time_length, obs_days, history, n_envs, device)
    _Env.__name__ = _Env.__qualname__ = name
    return _Env


Market_InvA_D1, Market_InvB_D1, Market_InvC_D1 = (_named(f"Market_Inv{i}_D1", i, False) for i in "ABC")
Market_InvA_Dx, Market_InvB_Dx, Market_InvC_Dx = (_named(f"Market_Inv{i}_Dx", i, True) for i in "ABC")
