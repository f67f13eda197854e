"""
Batched variant of the reference's multiplicative gym environments, under the
reference's own class names:

    Coin_InvA/B/C(n_gambles), Dice_InvA/B/C(n_gambles), GBM_InvA/B/C(n_gambles)
        (envs/coin_flip_envs.py, envs/dice_roll_envs.py, envs/gbm_envs.py)
    Dice_SH_INSURED(), Dice_SH_InvA/B/C()         (envs/dice_roll_sh_envs.py)

Each object steps `n_envs` independent copies in lock-step with one CUDA launch
(b200_menv_step).  With the default `n_envs=1` the interface is the reference's:
`reset() -> state`, `step(action) -> (next_state, reward, [done, learn_done],
risk)` with NumPy fp64 arrays / Python scalars, plus `observation_space`,
`action_space` (`shape, high, low, sample()`) and `reward_range`, so the object can
stand where `scripts/rl_multiplicative.py:68` builds the reference env.  With
`n_envs > 1` the same methods take / return CUDA fp64 tensors with a leading batch
dimension (done: bool [E,2]) and `reset(mask)` restarts only the masked copies -
the natural client is the 100 lock-step evaluation episodes of
tools/eval_episodes.py:233-263.

Differences from the reference, on purpose: outputs are fresh arrays (the
reference returns aliased internal buffers, envs/coin_flip_envs.py:188-216);
returns are drawn on the device from an explicit (seed, draw counter) Philox
stream instead of the unseeded global np.random state, or injected with
`step(action, returns=...)`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import EnvDesc, check, lib, ptr, require_cuda, stream_ptr

MAX_VALUE = 1e18
INITIAL_VALUE = 1e4
MIN_VALUE = max(1e-2 * INITIAL_VALUE, 1)
MAX_ABS_ACTION = 0.99
MAX_RETURN = 1e10
MIN_WEIGHT = 1e-5

_FAMILY = {
    "coin": dict(id=_lib.ENV_COIN, returns=(0.5, -0.4, 0.0), probs=(0.5, 0.5, 0.0), eta=1 / 0.5,
                 min_reward=1e-3, min_return=-0.9),
    "dice": dict(id=_lib.ENV_DICE, returns=(0.5, -0.5, 0.05), probs=(1 / 6, 1 / 6, 1 - (1 / 6 + 1 / 6)),
                 eta=1 / 0.5, min_reward=1e-3, min_return=-0.9),
    "gbm": dict(id=_lib.ENV_GBM, returns=(0.0, 0.0, 0.0), probs=(0.0, 0.0, 0.0), eta=5.0, min_reward=1e-3,
                min_return=math.log(0.1), drift=0.0540025395205692, vol=0.1897916175617430),
    "dice_sh": dict(id=_lib.ENV_DICE_SH, returns=(0.5, -0.5, 0.05), probs=(1 / 6, 1 / 6, 1 - (1 / 6 + 1 / 6)),
                    eta=1 / 0.5, min_reward=1e-6, min_return=-0.99, sh=(max(-1, -0.99), 5.0, max(-1, -0.99)),
                    i_eta=(-1 - 5) / (-0.5 - 5), sh_eta=1.0),
}
_INVESTOR = {"A": _lib.INV_A, "B": _lib.INV_B, "C": _lib.INV_C, "I": _lib.INV_INSURED}


class Box:
    """The slice of gym.spaces.Box the reference's loops touch."""

    def __init__(self, low, high, shape, dtype=np.float64):
        self.low = np.full(shape, low, dtype)
        self.high = np.full(shape, high, dtype)
        self.shape = shape
        self.dtype = dtype

    def sample(self):
        return np.random.uniform(self.low, self.high)


class BatchedMultiplicativeEnv:
    def __init__(self, family: str, investor: str, n_gambles: int = 1, n_envs: int = 1, seed: int = 0,
                 device="cuda"):
        require_cuda()
        f = _FAMILY[family]
        self.family, self.investor = family, investor
        self.n_gambles = 1 if family == "dice_sh" else int(n_gambles)
        self.n_envs = int(n_envs)
        self.device = torch.device(device)
        d = EnvDesc()
        d.family, d.investor, d.n_gambles = f["id"], _INVESTOR[investor], self.n_gambles
        d.stop_abs = 1 if (family == "coin" and investor == "B") else 0
        d.max_value, d.initial_value, d.min_value = MAX_VALUE, INITIAL_VALUE, MIN_VALUE
        d.max_abs_action, d.min_reward, d.min_return = MAX_ABS_ACTION, f["min_reward"], f["min_return"]
        d.max_return, d.min_weight, d.lev_factor = MAX_RETURN, MIN_WEIGHT, f["eta"]
        for i in range(3):
            d.returns[i], d.probs[i] = f["returns"][i], f["probs"][i]
            d.sh_returns[i] = f.get("sh", (0.0, 0.0, 0.0))[i]
        d.i_lev_factor, d.sh_lev_factor = f.get("i_eta", 0.0), f.get("sh_eta", 0.0)
        if family == "gbm":
            d.log_mean, d.vol = f["drift"] - f["vol"] ** 2 / 2, f["vol"]
        d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._d = d
        s, a, r = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.b200_menv_dims(C.byref(d), C.byref(s), C.byref(a), C.byref(r)))
        self.state_dim, self.action_dim, self.risk_dim = s.value, a.value, r.value
        self.reward_range = (f["min_reward"], np.inf)
        self.observation_space = Box(-np.inf, np.inf, (self.state_dim,))
        self.action_space = Box(-MAX_ABS_ACTION, MAX_ABS_ACTION, (self.action_dim,))
        e = self.n_envs
        with torch.cuda.device(self.device):
            self._wealth = torch.empty(e, dtype=torch.float64, device=self.device)
            self._time = torch.empty(e, dtype=torch.int32, device=self.device)
        self._draws = 0
        self.reset()

    # -- reference-style accessors (n_envs == 1)
    @property
    def wealth(self):
        return float(self._wealth[0]) if self.n_envs == 1 else self._wealth

    @property
    def time(self):
        return int(self._time[0]) if self.n_envs == 1 else self._time

    def reset(self, mask: Optional[torch.Tensor] = None):
        e = self.n_envs
        with torch.cuda.device(self.device):
            state = torch.empty((e, self.state_dim), dtype=torch.float64, device=self.device)
            m = None
            if mask is not None:
                m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
                # rows that are not reset keep no meaningful state here: fill with nan
                state.fill_(float("nan"))
            check(lib.b200_menv_reset(C.byref(self._d), e, ptr(self._wealth), ptr(self._time), ptr(state), ptr(m),
                                      stream_ptr()))
        return state[0].cpu().numpy() if e == 1 else state

    def step(self, action, returns=None):
        e = self.n_envs
        dev = self.device
        with torch.cuda.device(dev):
            a = torch.as_tensor(np.asarray(action, dtype=np.float64) if not isinstance(action, torch.Tensor) else action)
            a = a.to(device=dev, dtype=torch.float64).reshape(e, self.action_dim).contiguous()
            r = None
            if returns is not None:
                r = torch.as_tensor(np.asarray(returns, dtype=np.float64) if not isinstance(returns, torch.Tensor)
                                    else returns).to(device=dev, dtype=torch.float64)
                r = r.reshape(e, self.n_gambles).contiguous()
            ns = torch.empty((e, self.state_dim), dtype=torch.float64, device=dev)
            rew = torch.empty(e, dtype=torch.float64, device=dev)
            done = torch.empty((e, 2), dtype=torch.uint8, device=dev)
            risk = torch.empty((e, self.risk_dim), dtype=torch.float64, device=dev)
            check(lib.b200_menv_step(C.byref(self._d), e, ptr(self._wealth), ptr(self._time), ptr(a), ptr(r),
                                     C.c_uint64(self._draws), ptr(ns), ptr(rew), ptr(done), ptr(risk), stream_ptr()))
            self._draws += 1
        if e == 1:
            d = done.cpu().numpy()[0]
            return ns[0].cpu().numpy(), float(rew[0]), [bool(d[0]), bool(d[1])], risk[0].cpu().numpy()
        return ns, rew, done.bool(), risk


def _make(family, investor, takes_n):
    if takes_n:
        class _Env(BatchedMultiplicativeEnv):
            def __init__(self, n_gambles: int, n_envs: int = 1, seed: int = 0, device="cuda"):
                super().__init__(family, investor, n_gambles, n_envs, seed, device)
    else:
        class _Env(BatchedMultiplicativeEnv):
            def __init__(self, n_envs: int = 1, seed: int = 0, device="cuda"):
                super().__init__(family, investor, 1, n_envs, seed, device)
    return _Env


def _named(name, family, investor, takes_n=True):
    cls = _make(family, investor, takes_n)
    cls.__name__ = cls.__qualname__ = name
    return cls


Coin_InvA, Coin_InvB, Coin_InvC = (_named(f"Coin_Inv{i}", "coin", i) for i in "ABC")
Dice_InvA, Dice_InvB, Dice_InvC = (_named(f"Dice_Inv{i}", "dice", i) for i in "ABC")
GBM_InvA, GBM_InvB, GBM_InvC = (_named(f"GBM_Inv{i}", "gbm", i) for i in "ABC")
Dice_SH_INSURED = _named("Dice_SH_INSURED", "dice_sh", "I", takes_n=False)
Dice_SH_InvA, Dice_SH_InvB, Dice_SH_InvC = (_named(f"Dice_SH_Inv{i}", "dice_sh", i, takes_n=False) for i in "ABC")
