"""
`ReplayBufferTorch` - the reference's experience replay (tools/replay_torch.py;
NumPy twin tools/replay.py) on the B200 engine, under the reference's own class
name, constructor (`inputs` dictionary, :57-115) and methods:

    store_exp(state, action, reward, next_state, done)           (:167-197)
    sample_exp() -> (states, actions, rewards, next_states, dones, eff_length)   (:360-412)

plus the attributes the agents read (`mem_idx`, `mem_size`, `batch_size`, the
`*_memory` tensors; algos/algo_sac.py:106-110,190,259,380).

What changes underneath (rlmd_b200/csrc/replay.cu):

* store_exp is one kernel launch with the transition as its parameter (the
  reference issues five host->device copies and, for multi_steps > 1, two
  O(mem_size) scans per store, :147,:154);
* the per-sample episode search and the Python n-step loops of sample_exp are
  one gather kernel reading a per-slot episode-start array;
* indices are drawn on the device (distinct, uniform - what
  `randperm(max_mem)[:batch]` yields, :383 - from an explicit Philox stream
  `(seed, draw counter)`), or injected with `sample_exp(batch=...)`;
* `store_batch` appends many transitions at once (device tensors, e.g. the
  output of a batched env) and `sample_many(k)` draws k mini-batches per launch.

Outputs are fresh tensors (the reference returns its aliased `multi_*` buffers
for multi_steps > 1, :404-410).  As in the reference the multi-step bookkeeping
assumes an append-only buffer (`buffer >= n_cumsteps`,
tests/test_input_agent.py:280-291): a multi-step buffer refuses to wrap.
There is no CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import ReplayDesc, check, lib, ptr, require_cuda, stream_ptr


class ReplayBufferTorch:
    def __init__(self, inputs: dict, seed: int = 0) -> None:
        require_cuda()
        gpu = inputs.get("gpu", "cuda:0")
        self.device = torch.device(gpu if str(gpu).startswith("cuda") else "cuda:0")
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()

        self.input_dims = int(sum(inputs["input_dims"]))
        self.num_actions = int(inputs["num_actions"])
        self.batch_size = int(inputs["mini_batch_size"])
        self.gamma = inputs["discount"]
        self.multi_steps = torch.tensor(int(inputs["multi_steps"]), device=self.device)
        self._n = int(inputs["multi_steps"])
        if not 1 <= self._n <= _lib.REPLAY_MAX_STEPS:
            raise ValueError(f"multi_steps must be in 1..{_lib.REPLAY_MAX_STEPS}")
        self.r_abs_zero = -np.inf if inputs["r_abs_zero"] is None else inputs["r_abs_zero"]
        self.dyna = str(inputs["dynamics"])
        if int(inputs["buffer"]) <= int(inputs["n_cumsteps"]):
            self.mem_size = int(inputs["buffer"])
        else:
            self.mem_size = int(inputs["n_cumsteps"])
        self.mem_idx = 0
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._draws = 0

        dev, m = self.device, self.mem_size
        with torch.cuda.device(dev):
            self.state_memory = torch.zeros((m, self.input_dims), dtype=torch.float32, device=dev)
            self.action_memory = torch.zeros((m, self.num_actions), dtype=torch.float32, device=dev)
            self.reward_memory = torch.zeros((m,), dtype=torch.float32, device=dev)
            self.next_state_memory = torch.zeros((m, self.input_dims), dtype=torch.float32, device=dev)
            self.terminal_memory = torch.zeros((m,), dtype=torch.bool, device=dev)
            self.eff_length = torch.ones((1,), dtype=torch.int, device=dev)
            self.episode_start = torch.zeros((m,), dtype=torch.int32, device=dev)
            self.header = torch.zeros((8,), dtype=torch.int64, device=dev)

        d = ReplayDesc()
        d.mem_size, d.state_dim, d.action_dim = m, self.input_dims, self.num_actions
        d.state_memory, d.action_memory = self.state_memory.data_ptr(), self.action_memory.data_ptr()
        d.reward_memory, d.next_state_memory = self.reward_memory.data_ptr(), self.next_state_memory.data_ptr()
        d.terminal_memory, d.episode_start = self.terminal_memory.data_ptr(), self.episode_start.data_ptr()
        d.header = self.header.data_ptr()
        self._desc = d
        # fl32(gamma**t), the factor the reference forms per term (:300)
        self._gamma_pow = (C.c_float * _lib.REPLAY_MAX_STEPS)(
            *[float(np.float32(float(self.gamma) ** t)) for t in range(_lib.REPLAY_MAX_STEPS)])
        w = 2 * self.input_dims + self.num_actions
        self._row = np.zeros(w, dtype=np.float64)
        self._row_ptrs = tuple(
            self._row[a:].ctypes.data_as(C.POINTER(C.c_double))
            for a in (0, self.input_dims, self.input_dims + self.num_actions))
        self._host_row_ok = w + 1 <= 480

    # ------------------------------------------------------------------ store
    def _check_room(self, count: int) -> None:
        if self._n > 1 and self.mem_idx + count > self.mem_size:
            raise RuntimeError(
                "multi-step replay is append-only (the reference asserts buffer >= n_cumsteps): "
                f"{self.mem_idx + count} transitions do not fit mem_size {self.mem_size}")

    def store_exp(self, state, action, reward: float, next_state, done: bool) -> None:
        """tools/replay_torch.py:167-197."""
        self._check_room(1)
        s, a = self.input_dims, self.num_actions
        if not self._host_row_ok:
            dev = self.device
            self.store_batch(torch.as_tensor(np.asarray(state, np.float64).reshape(1, s), device=dev),
                             torch.as_tensor(np.asarray(action, np.float64).reshape(1, a), device=dev),
                             torch.as_tensor(np.asarray([reward], np.float64), device=dev),
                             torch.as_tensor(np.asarray(next_state, np.float64).reshape(1, s), device=dev),
                             torch.as_tensor(np.asarray([bool(done)]), device=dev))
            return
        row = self._row
        row[:s] = np.asarray(state, dtype=np.float64).reshape(-1)
        row[s:s + a] = np.asarray(action, dtype=np.float64).reshape(-1)
        row[s + a:] = np.asarray(next_state, dtype=np.float64).reshape(-1)
        with torch.cuda.device(self.device):
            check(lib.b200_replay_store_host(C.byref(self._desc), self._row_ptrs[0], self._row_ptrs[1], float(reward),
                                             self._row_ptrs[2], int(bool(done)), self.mem_idx,
                                             float(self.r_abs_zero), stream_ptr()))
        self.mem_idx += 1

    def store_batch(self, states, actions, rewards, next_states, dones) -> None:
        """Appends `count` transitions in order (device tensors, fp64 or fp32; dones bool / uint8)."""
        dev = self.device
        dt = states.dtype
        if dt not in (torch.float64, torch.float32):
            raise ValueError("store_batch takes float64 or float32 tensors")
        count = int(rewards.shape[0])
        self._check_room(count)
        st = states.to(device=dev, dtype=dt).reshape(count, self.input_dims).contiguous()
        ac = actions.to(device=dev, dtype=dt).reshape(count, self.num_actions).contiguous()
        rw = rewards.to(device=dev, dtype=dt).reshape(count).contiguous()
        ns = next_states.to(device=dev, dtype=dt).reshape(count, self.input_dims).contiguous()
        dn = dones.to(device=dev).reshape(count).to(torch.uint8).contiguous()
        with torch.cuda.device(dev):
            check(lib.b200_replay_store(C.byref(self._desc), ptr(st), ptr(ac), ptr(rw), ptr(ns), ptr(dn), count,
                                        int(dt == torch.float64), self.mem_idx, float(self.r_abs_zero), stream_ptr()))
        self.mem_idx += count

    # ----------------------------------------------------------------- sample
    def _sample(self, k: int, b: int, batch: Optional[torch.Tensor]):
        dev = self.device
        max_mem = min(self.mem_idx, self.mem_size)
        # entering a device context costs several microseconds of a ~35 us call: only when it is needed
        ctx = contextlib.nullcontext() if torch.cuda.current_device() == self._dev_index else torch.cuda.device(dev)
        with ctx:
            n = k * b
            out_s = torch.empty((n, self.input_dims), dtype=torch.float32, device=dev)
            out_a = torch.empty((n, self.num_actions), dtype=torch.float32, device=dev)
            out_r = torch.empty((n,), dtype=torch.float32, device=dev)
            out_s2 = torch.empty((n, self.input_dims), dtype=torch.float32, device=dev)
            out_d = torch.empty((n,), dtype=torch.bool, device=dev)
            out_e = torch.empty((n,), dtype=torch.int64, device=dev)
            if batch is None:
                idx = None
                out_i = torch.empty((n,), dtype=torch.int64, device=dev)
                self._draws += 1
            else:
                idx = torch.as_tensor(batch, device=dev).to(torch.int64).reshape(n).contiguous()
                out_i = idx
            check(lib.b200_replay_sample(C.byref(self._desc), ptr(idx), k, b, max_mem, self._n, self._gamma_pow,
                                         int(self.dyna == "A"), self.seed, self._draws, ptr(out_i), ptr(out_s),
                                         ptr(out_a), ptr(out_r), ptr(out_s2), ptr(out_d), ptr(out_e), stream_ptr()))
        return out_i, out_s, out_a, out_r, out_s2, out_d, out_e

    def sample_exp(self, batch=None):
        """
        tools/replay_torch.py:360-412.  `batch` (optional) = the slots to sample,
        standing where the reference draws `randperm(max_mem)[:batch_size]`.
        """
        max_mem = min(self.mem_idx, self.mem_size)
        if batch is None:
            b = min(self.batch_size, max_mem)      # randperm(max_mem)[:B] is shorter on a young buffer
            if self._n > 1 and b < self.batch_size:
                raise IndexError("multi-step sampling needs at least mini_batch_size stored transitions "
                                 "(the reference's per-sample loop runs over batch_size, :347-352)")
        else:
            b = int(torch.as_tensor(batch).numel())
        _, s, a, r, s2, d, e = self._sample(1, b, batch)
        self.last_batch = _
        eff = e if self._n > 1 else self.eff_length[0]
        return s, a, r, s2, d, eff

    def capture_sampler(self, k: int = 1):
        """
        `sample_exp` as a CUDA graph: records b200_replay_sample_counted (draw on the device, `filled` from
        the device header, draw index from a device counter) ONCE into caller-owned static output tensors and
        returns a callable; each call replays the graph - no allocation, no foreign-function call - and
        returns the reference's 6-tuple (states, actions, rewards, next_states, dones, eff_length) as views
        of the static tensors (overwritten by the next call; `self.last_batch` holds the drawn slots).
        Transitions stored after the capture are seen (the header lives on the device).  Needs
        mini_batch_size stored transitions, like the reference's multi-step sampler.  k > 1: k mini-batches
        per replay, rows [k * batch_size, ...].
        """
        dev, b, n = self.device, self.batch_size, int(k) * self.batch_size
        if min(self.mem_idx, self.mem_size) < b:
            raise IndexError("fewer stored transitions than mini_batch_size")
        with torch.cuda.device(dev):
            out = dict(idx=torch.empty((n,), dtype=torch.int64, device=dev),
                       s=torch.empty((n, self.input_dims), dtype=torch.float32, device=dev),
                       a=torch.empty((n, self.num_actions), dtype=torch.float32, device=dev),
                       r=torch.empty((n,), dtype=torch.float32, device=dev),
                       s2=torch.empty((n, self.input_dims), dtype=torch.float32, device=dev),
                       d=torch.empty((n,), dtype=torch.bool, device=dev),
                       e=torch.empty((n,), dtype=torch.int64, device=dev))
            # its own stream of draw indices, far from the ones sample_exp() uses
            counter = torch.tensor([(1 << 40) + self._draws, 0], dtype=torch.int64, device=dev)

            def launch():
                check(lib.b200_replay_sample_counted(C.byref(self._desc), int(k), b, self._n, self._gamma_pow,
                                                     int(self.dyna == "A"), self.seed, ptr(counter), ptr(out["idx"]),
                                                     ptr(out["s"]), ptr(out["a"]), ptr(out["r"]), ptr(out["s2"]),
                                                     ptr(out["d"]), ptr(out["e"]), stream_ptr()))

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                launch()                      # warm-up outside the capture (module loading, function attributes)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                launch()
        eff = out["e"] if self._n > 1 else self.eff_length[0]
        result = (out["s"], out["a"], out["r"], out["s2"], out["d"], eff)
        self._sampler = (g, out, counter)

        def replay():
            g.replay()
            self.last_batch = out["idx"]
            return result

        return replay

    def sample_many(self, k: int, batches=None):
        """k mini-batches in one launch: tensors with a leading [k, batch_size] shape (+ the slots)."""
        b = self.batch_size
        if min(self.mem_idx, self.mem_size) < b:
            raise IndexError("fewer stored transitions than mini_batch_size")
        i, s, a, r, s2, d, e = self._sample(int(k), b, batches)
        return (i.view(k, b), s.view(k, b, -1), a.view(k, b, -1), r.view(k, b), s2.view(k, b, -1), d.view(k, b),
                e.view(k, b))
