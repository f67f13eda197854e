"""
Fused collector and batched evaluation: the two
callers that sit on either side of the env step and the replay buffer in the
reference,

* the interaction loop of scripts/rl_multiplicative.py:185-273
      next_state, reward, env_done, risk = env.step(action)
      agent.store_transistion(state, action, reward, next_state, learn_done)
      state = next_state                      # env.reset() after a done step
      ... agent.learn() -> replay.sample_exp()
* `eval_multiplicative` of tools/eval_episodes.py:176-315 (n_eval episodes under
  one constant action each; mean / median / 5th percentile / MAD / std of the
  growth reward, the valuation and the episode length),

as device-side pipelines without a host round trip per step.  `Collector.step`
is ONE kernel (env step + append + observation carry) for `n_envs` environments,
each writing its own lane of the replay memory (a lane is a reference replay
buffer, its j-th transition at slot `j * n_envs + lane`; `n_envs = 1` is the
reference's single stream); `Collector.sample` is
the n-step gather over the lanes; `Collector.capture` records step + sample in
a CUDA graph.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from datetime import datetime
from typing import Optional

import numpy as np
import torch

from . import _lib, envs
from ._lib import CollectDesc, ReplayDesc, check, lib, ptr, require_cuda, stream_ptr


class Collector:
    """
    env        a BatchedMultiplicativeEnv instance (any class of rlmd_b200.envs) - its
               description and seed are used, its own wealth/time buffers are not
    lane_len   transitions kept per environment (the reference's `buffer` for n_envs = 1)
    inputs     the reference's `inputs` keys the replay buffer reads: mini_batch_size,
               discount, multi_steps, r_abs_zero, dynamics (tools/replay_torch.py:64-82)
    """

    def __init__(self, env: "envs.BatchedMultiplicativeEnv", lane_len: int, inputs: dict, seed: int = 0):
        require_cuda()
        self.env = env
        self.device = env.device
        self.n_envs, self.lane_len = int(env.n_envs), int(lane_len)
        self.state_dim, self.action_dim, self.risk_dim = env.state_dim, env.action_dim, env.risk_dim
        self.batch_size = int(inputs["mini_batch_size"])
        self.gamma = inputs["discount"]
        self._n = int(inputs["multi_steps"])
        if not 1 <= self._n <= _lib.REPLAY_MAX_STEPS:
            raise ValueError(f"multi_steps must be in 1..{_lib.REPLAY_MAX_STEPS}")
        self.dyna = str(inputs["dynamics"])
        self.r_abs_zero = -np.inf if inputs.get("r_abs_zero") is None else float(inputs["r_abs_zero"])
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        e, m, dev = self.n_envs, self.n_envs * self.lane_len, self.device
        if m >= 1 << 31:
            raise ValueError("n_envs * lane_len must be < 2^31")
        self.mem_size = m
        with torch.cuda.device(dev):
            self.state_memory = torch.zeros((m, self.state_dim), dtype=torch.float32, device=dev)
            self.action_memory = torch.zeros((m, self.action_dim), dtype=torch.float32, device=dev)
            self.reward_memory = torch.zeros((m,), dtype=torch.float32, device=dev)
            self.next_state_memory = torch.zeros((m, self.state_dim), dtype=torch.float32, device=dev)
            self.terminal_memory = torch.zeros((m,), dtype=torch.bool, device=dev)
            self.episode_start = torch.zeros((m,), dtype=torch.int32, device=dev)
            self.header = torch.zeros((e, 8), dtype=torch.int64, device=dev)
            self.wealth = torch.empty(e, dtype=torch.float64, device=dev)
            self.time = torch.empty(e, dtype=torch.int32, device=dev)
            self.state = torch.empty((e, self.state_dim), dtype=torch.float64, device=dev)   # the observation
            self.counter = torch.zeros(2, dtype=torch.int64, device=dev)
            self.reward = torch.empty(e, dtype=torch.float64, device=dev)
            self.done = torch.empty((e, 2), dtype=torch.uint8, device=dev)
            self.risk = torch.empty((e, self.risk_dim), dtype=torch.float64, device=dev)
        r = ReplayDesc()
        r.mem_size, r.state_dim, r.action_dim = m, self.state_dim, self.action_dim
        r.state_memory, r.action_memory = self.state_memory.data_ptr(), self.action_memory.data_ptr()
        r.reward_memory, r.next_state_memory = self.reward_memory.data_ptr(), self.next_state_memory.data_ptr()
        r.terminal_memory, r.episode_start = self.terminal_memory.data_ptr(), self.episode_start.data_ptr()
        r.header = self.header.data_ptr()
        d = CollectDesc()
        d.env, d.replay = env._d, r
        d.n_envs, d.lane_len, d.reward_floor = e, self.lane_len, self.r_abs_zero
        self._desc = d
        self._gamma_pow = (C.c_float * _lib.REPLAY_MAX_STEPS)(
            *[float(np.float32(float(self.gamma) ** t)) for t in range(_lib.REPLAY_MAX_STEPS)])
        self.steps = 0
        self._graph = None
        self.reset()

    # ------------------------------------------------------------------ loop
    def reset(self) -> torch.Tensor:
        with torch.cuda.device(self.device):
            check(lib.b200_collect_reset(C.byref(self._desc), ptr(self.wealth), ptr(self.time), ptr(self.state),
                                         ptr(self.counter), stream_ptr()))
        self.steps = 0
        return self.state

    def step(self, action: torch.Tensor, returns: Optional[torch.Tensor] = None) -> torch.Tensor:
        """
        One env step + append for every environment.  action: float64 CUDA [n_envs, A]
        (read in place - the agent's output buffer); returns: injected returns
        [n_envs, n_gambles] or None (Philox).  Returns `self.state` (the next
        observation, updated in place); `self.reward / done / risk` hold the step's outputs.
        """
        e = self.n_envs
        if self._n > 1 and self.steps + 1 > self.lane_len:
            raise RuntimeError("multi-step replay is append-only (the reference asserts buffer >= n_cumsteps)")
        if not (action.is_cuda and action.dtype == torch.float64 and action.is_contiguous()
                and tuple(action.shape) == (e, self.action_dim)):
            raise ValueError("action must be a contiguous float64 CUDA tensor [n_envs, action_dim]")
        if returns is not None and not (returns.is_cuda and returns.dtype == torch.float64 and returns.is_contiguous()
                                        and returns.numel() == e * self.env.n_gambles):
            raise ValueError("returns must be a contiguous float64 CUDA tensor [n_envs, n_gambles]")
        with torch.cuda.device(self.device):
            check(lib.b200_collect_step(C.byref(self._desc), ptr(self.wealth), ptr(self.time), ptr(self.state),
                                        ptr(action), ptr(returns), ptr(self.counter), ptr(self.reward),
                                        ptr(self.done), ptr(self.risk), stream_ptr()))
        self.steps += 1
        return self.state

    def _alloc_batch(self, k: int):
        b, dev = self.batch_size, self.device
        n = k * b
        return dict(
            idx=torch.empty((k, b), dtype=torch.int64, device=dev),
            states=torch.empty((n, self.state_dim), dtype=torch.float32, device=dev),
            actions=torch.empty((n, self.action_dim), dtype=torch.float32, device=dev),
            rewards=torch.empty((n,), dtype=torch.float32, device=dev),
            next_states=torch.empty((n, self.state_dim), dtype=torch.float32, device=dev),
            dones=torch.empty((n,), dtype=torch.bool, device=dev),
            eff=torch.empty((n,), dtype=torch.int64, device=dev),
        )

    def _sample_into(self, out: dict, k: int, batch: Optional[torch.Tensor]) -> None:
        with torch.cuda.device(self.device):
            check(lib.b200_collect_sample(
                C.byref(self._desc), ptr(batch), k, self.batch_size, self._n, self._gamma_pow,
                1 if self.dyna == "A" else 0, C.c_uint64(self.seed), ptr(self.counter), ptr(out["idx"]),
                ptr(out["states"]), ptr(out["actions"]), ptr(out["rewards"]), ptr(out["next_states"]),
                ptr(out["dones"]), ptr(out["eff"]), stream_ptr()))

    def sample(self, k: int = 1, batch=None):
        """
        k mini-batches of the reference's `sample_exp` 6-tuple over all lanes
        (states, actions, rewards, next_states, dones, eff_length), shaped [k*B, ...].
        batch: optional int64 slots [k, B] (slot = local index * n_envs + lane).
        """
        if batch is not None:
            batch = torch.as_tensor(batch, device=self.device).to(torch.int64).reshape(k, self.batch_size).contiguous()
        elif self.steps * self.n_envs < self.batch_size:
            # fewer stored transitions than a mini-batch of distinct slots: the reference's randperm()[:B] then
            # indexes out of range (tools/replay_torch.py:383)
            raise IndexError(f"{self.steps * self.n_envs} transitions stored, mini-batch size {self.batch_size}")
        with torch.cuda.device(self.device):
            out = self._alloc_batch(k)
        self._sample_into(out, k, batch)
        self.last_batch = out["idx"] if batch is None else batch
        return out["states"], out["actions"], out["rewards"], out["next_states"], out["dones"], out["eff"]

    # ----------------------------------------------------------------- graph
    def capture(self, action: torch.Tensor, k: int = 1):
        """
        Records `step(action)` followed by `sample(k)` in a CUDA graph.  `action` is
        read in place at every replay (write the policy's output into it).  Returns
        a callable; each call replays the graph and returns the static output dict
        (keys: states, actions, rewards, next_states, dones, eff, idx).
        NOTE: capturing performs ONE real `step(action)` + sample as its warm-up (CUDA
        needs the kernels loaded before a capture): a transition is appended, the envs
        advance one step and both Philox counters move on - `self.steps` counts it.
        Until `batch_size` transitions are stored the drawn slots are -1 and the
        sampled rows are zero with eff = 0 (check `out["idx"]`).
        """
        dev = self.device
        with torch.cuda.device(dev):
            out = self._alloc_batch(k)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):      # warm-up outside the capture (lazy module loading, attributes)
                self.step(action)
                self._sample_into(out, k, None)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                check(lib.b200_collect_step(C.byref(self._desc), ptr(self.wealth), ptr(self.time), ptr(self.state),
                                            ptr(action), None, ptr(self.counter), ptr(self.reward), ptr(self.done),
                                            ptr(self.risk), stream_ptr()))
                self._sample_into(out, k, None)
        self._graph = g

        def replay():
            if self._n > 1 and self.steps + 1 > self.lane_len:
                raise RuntimeError("multi-step replay is append-only (the reference asserts buffer >= n_cumsteps)")
            g.replay()
            self.steps += 1
            return out

        return replay


# ------------------------------------------------------------------ evaluation
def rollout(env: "envs.BatchedMultiplicativeEnv", actions, max_steps: int, *, returns=None, draw_base: int = 0):
    """
    `n` evaluation episodes in one launch (tools/eval_episodes.py:233-262): episode i
    holds actions[i] constant from the reset state until done or `max_steps` steps.
    returns: optional injected returns [max_steps, n, n_gambles] (float64).
    -> (reward [n] f64, steps [n] i32, risk [n,R] f64, last_state [n,S] f64) on the GPU.
    """
    require_cuda()
    dev = env.device
    a = torch.as_tensor(actions, device=dev).to(torch.float64)
    a = a.reshape(-1, env.action_dim).contiguous()
    n = a.shape[0]
    r = None
    if returns is not None:
        r = torch.as_tensor(returns, device=dev).to(torch.float64).contiguous()
        if r.numel() != int(max_steps) * n * env.n_gambles:
            raise ValueError("returns must hold max_steps x n x n_gambles values")
    with torch.cuda.device(dev):
        reward = torch.zeros(n, dtype=torch.float64, device=dev)
        steps = torch.zeros(n, dtype=torch.int32, device=dev)
        risk = torch.zeros((n, env.risk_dim), dtype=torch.float64, device=dev)
        last = torch.zeros((n, env.state_dim), dtype=torch.float64, device=dev)
        check(lib.b200_menv_rollout(C.byref(env._d), n, ptr(a), ptr(r), C.c_uint64(int(draw_base)), int(max_steps),
                                    ptr(reward), ptr(steps), ptr(risk), ptr(last), stream_ptr()))
    return reward, steps, risk, last


def eval_summary(reward: np.ndarray, val: np.ndarray, lev: np.ndarray, step: np.ndarray) -> list:
    """The 15 statistics of tools/eval_episodes.py:264-302, in the reference's order."""
    pct = lambda x, q: np.percentile(x, q=q, method="median_unbiased")
    mean_reward, mean_val, mean_step = np.mean(reward), np.mean(val), np.mean(step)
    return [
        np.mean(lev) * 100,
        (mean_reward - 1) * 100, (pct(reward, 50) - 1) * 100, (pct(reward, 5) - 1) * 100,
        np.mean(np.abs(reward - mean_reward)) * 100, np.std(reward, ddof=0) * 100,
        mean_val, pct(val, 50), pct(val, 5), np.mean(np.abs(val - mean_val)),
        mean_step, pct(step, 50), pct(step, 5), np.mean(np.abs(step - mean_step)), np.std(step, ddof=0),
    ]


def eval_multiplicative(env: "envs.BatchedMultiplicativeEnv", run_action, n_eval: int, max_eval_steps: int, *,
                        returns=None, draw_base: int = 0, verbose: bool = True) -> dict:
    """
    The evaluation block of `eval_multiplicative` (tools/eval_episodes.py:231-399) for
    one policy action: `n_eval` lock-step episodes on the GPU, then the reference's
    summary (`stats`, and the per-episode columns it logs: reward, steps, risk).
    run_action: [A] (every episode starts from the same reset state, so the
    reference's deterministic `agent.eval_next_action(reset state)` is one vector)
    or [n_eval, A].
    """
    a = np.asarray(run_action.detach().cpu() if isinstance(run_action, torch.Tensor) else run_action, dtype=np.float64)
    if a.ndim == 1:
        a = np.tile(a, (int(n_eval), 1))
    t0 = datetime.now()
    reward, steps, risk, _ = rollout(env, a, int(max_eval_steps), returns=returns, draw_base=draw_base)
    reward, steps, risk = reward.cpu().numpy(), steps.cpu().numpy().astype(np.float64), risk.cpu().numpy()
    secs = max((datetime.now() - t0).total_seconds(), 1e-9)
    stats = eval_summary(reward, risk[:, 1], risk[:, 3], steps)
    out = dict(reward=reward, steps=steps, risk=risk, stats=stats, steps_sec=float(steps.sum() / secs))
    if verbose:
        head = "{} Summary {:1.0f}/s ".format(datetime.now().strftime("%d %H:%M:%S"), out["steps_sec"])
        if env.investor in ("A", "I"):
            lev_part = "l% {:1.0f} ".format(stats[0])
        elif env.investor == "B":
            lev_part = "l%/s% {:1.0f}/{:1.0f} ".format(stats[0], np.mean(risk[:, 4]) * 100)
        else:
            lev_part = "l%/s%/r% {:1.0f}/{:1.0f}/{:1.0f} ".format(stats[0], np.mean(risk[:, 4]) * 100,
                                                                   np.mean(risk[:, 5]) * 100)
        print(head + lev_part + "mean/med/95/mad/std: g% {:1.1f}/{:1.1f}/{:1.1f}/{:1.0f}/{:1.0f} "
              "V$ {:1.1E}/{:1.1E}/{:1.1E}/{:1.0E} st {:1.0f}/{:1.0f}/{:1.0f}/{:1.0f}/{:1.0f}".format(*stats[1:]))
    return out
