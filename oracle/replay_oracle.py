"""
TEST INFRASTRUCTURE ONLY - CPU restatement of the reference replay buffer
(tools/replay_torch.py; tools/replay.py is the same algorithm on fp64 NumPy).

Nothing in rlmd_b200/ may import this module; it is the checker for the CUDA
replay path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).

Parity pin: tests/test_oracle_replay.py checks `ReplayOracle` against
tests/golden/replay_*.npz, which tests/golden/gen_golden_more.py:gen_replay
produced by driving the UNMODIFIED reference class through store_exp /
sample_exp with its `randperm` answering pre-drawn batches.

`ReplayOracle` deliberately follows the reference's own bookkeeping (lists of
episode slices, the try/except ladder) instead of the closed form of SURVEY.md
App. D; `closed_form_history` is that closed form, kept next to it so that the
two can be tested against each other - the CUDA kernel implements the latter.
"""
import numpy as np


class ReplayOracle:
    """tools/replay_torch.py:57-412 on NumPy fp32 arrays (append-only regime)."""

    def __init__(self, inputs: dict):
        # tools/replay_torch.py:64-82
        self.input_dims = int(sum(inputs["input_dims"]))
        self.num_actions = int(inputs["num_actions"])
        self.batch_size = int(inputs["mini_batch_size"])
        self.gamma = inputs["discount"]
        self.multi_steps = int(inputs["multi_steps"])
        self.r_abs_zero = -np.inf if inputs["r_abs_zero"] is None else inputs["r_abs_zero"]
        self.dyna = str(inputs["dynamics"])
        self.mem_size = int(min(int(inputs["buffer"]), int(inputs["n_cumsteps"])))
        self.mem_idx = 0
        # :86-100 (T.empty there; zeros here = the state the fixtures start from)
        self.state_memory = np.zeros((self.mem_size, self.input_dims), np.float32)
        self.action_memory = np.zeros((self.mem_size, self.num_actions), np.float32)
        self.reward_memory = np.zeros((self.mem_size,), np.float32)
        self.next_state_memory = np.zeros((self.mem_size, self.input_dims), np.float32)
        self.terminal_memory = np.zeros((self.mem_size,), bool)
        # :103-106; episode histories are (start, stop) slices of the memories,
        # which is what the reference's tensor views are
        self.epis_idx = [float("nan")]
        self.epis_slices = []
        self.n_terminal = 0

    # ---- tools/replay_torch.py:117-165
    def _episode_history(self, idx: int, done: bool) -> None:
        self.epis_idx[-1] = idx
        if len(self.epis_idx) >= 2:
            current = (self.epis_idx[-2] + 1, idx + 1)      # :130-133
        else:
            current = (0, idx + 1)                          # :137-139
        if self.n_terminal == 0 and done is not True:        # :147-151 first episode, live
            self.epis_idx = [idx + 1]
            self.epis_slices = [current]
        if self.n_terminal == 1 and done is True:            # :154-158 first episode ends
            self.epis_idx = [idx]
            self.epis_slices = []
        if done is True:                                     # :161-165
            self.epis_idx.append(idx + 1)
            self.epis_slices.append(current)

    # ---- tools/replay_torch.py:167-197
    def store_exp(self, state, action, reward, next_state, done) -> None:
        idx = self.mem_idx % self.mem_size
        self.state_memory[idx] = np.asarray(state, dtype=np.float32)
        self.action_memory[idx] = np.asarray(action, dtype=np.float32)
        self.reward_memory[idx] = np.float32(max(reward, self.r_abs_zero))
        self.next_state_memory[idx] = np.asarray(next_state, dtype=np.float32)
        self.terminal_memory[idx] = bool(done)
        self.n_terminal += int(bool(done))
        if self.multi_steps > 1:
            self._episode_history(idx, done)
        self.mem_idx += 1

    # ---- tools/replay_torch.py:199-247: (start, length) of the history of `step`
    def _construct_history(self, step: int):
        eh = self.epis_idx
        if step > eh[0]:                                     # :216-219
            sample_idx = max(j for j, e in enumerate(eh) if step - e > 0) + 1
            n_rewards = step - eh[sample_idx - 1]
        else:                                                # :220-221
            sample_idx, n_rewards = 0, step
        if sample_idx < len(self.epis_slices):               # :224-227
            lo, hi = self.epis_slices[sample_idx]
        elif self.epis_slices:                               # :229-233 falls back to episode 0
            lo, hi = self.epis_slices[0]
        else:                                                # :241-245 zero history
            return None
        return lo, min(hi - lo, n_rewards + 1)

    # ---- tools/replay_torch.py:273-310 + :336-345
    def _multi_step(self, hist):
        if hist is None:
            eff = 1
            r = np.float32(0.0) if self.dyna == "A" else np.float32(1.0)
            return r, np.zeros(self.input_dims, np.float32), np.zeros(self.num_actions, np.float32), eff
        lo, length = hist
        eff = min(length, self.multi_steps)
        first = lo + length - eff
        # gamma**t is a Python double; times a 0-dim fp32 tensor it is rounded to fp32
        # first and multiplied in fp32 (:300); sum / prod of <= n-1 terms in fp32
        terms = [np.float32(np.float32(self.gamma ** t) * self.reward_memory[first + t]) for t in range(eff - 1)]
        if self.dyna == "A":
            acc = np.float32(0.0)
            for x in terms:
                acc = np.float32(acc + x)
        else:
            acc = np.float32(1.0)
            for x in terms:
                acc = np.float32(acc * x)
        return acc, self.next_state_memory[first].copy(), self.action_memory[first].copy(), eff

    # ---- tools/replay_torch.py:360-412 with the index draw injected
    def sample_exp(self, batch):
        batch = np.asarray(batch, dtype=np.int64)
        states = self.state_memory[batch].copy()
        actions = self.action_memory[batch].copy()
        rewards = self.reward_memory[batch].copy()
        next_states = self.next_state_memory[batch].copy()
        dones = self.terminal_memory[batch].copy()
        eff = np.ones((1,), np.int64)
        if self.multi_steps > 1:
            eff = np.empty(len(batch), np.int64)
            for k, step in enumerate(batch):
                rewards[k], states[k], actions[k], eff[k] = self._multi_step(self._construct_history(int(step)))
        return states, actions, rewards, next_states, dones, eff


def closed_form_history(step: int, terminal: np.ndarray, filled: int):
    """
    SURVEY.md App. D: (start, length) of the history of `step` from the done flags
    alone (append-only buffer holding `filled` transitions) - what the CUDA
    kernel evaluates from its per-slot episode-start array and three scalars.
    """
    ends = np.flatnonzero(terminal[:filled])
    m = len(ends)
    if m == 0 or step <= ends[0]:
        return 0, step + 1
    if step <= ends[-1]:
        j = int(np.searchsorted(ends, step))          # ends[j-1] < step <= ends[j]
        start = int(ends[j - 1]) + 1
        return start, (step - start + 1) + (0 if terminal[step] else 1)
    n_rewards = step - int(ends[-1])
    return 0, min(n_rewards + 1, int(ends[0]) + 1)


def draw_unique(seed: int, draw_index: int, batch_id: int, filled: int, batch: int) -> np.ndarray:
    """
    The engine's on-device index draw (rlmd_b200/csrc/replay.cu:replay_draw_kernel),
    word for word: `batch` DISTINCT uniform slots in [0, filled) - the set
    `randperm(max_mem)[:batch]` yields (tools/replay_torch.py:383), from an
    explicit Philox stream.  Round r: every pending position t draws
        v = (philox(t, r, batch_id, lo32(draw_index) ^ TAG; key).xy * filled) >> 64
    and the claim with the smallest (round, t) owns v; the others redraw.
    """
    from oracle import philox_oracle as po

    if filled < batch or filled <= 0:
        return np.full(batch, -1, dtype=np.int64)
    k0 = (seed & 0xFFFFFFFF) ^ ((draw_index >> 32) & 0xFFFFFFFF)
    k1 = (seed >> 32) & 0xFFFFFFFF
    out = np.full(batch, -1, dtype=np.int64)
    owner = {}
    pending = list(range(batch))
    rnd = 0
    while pending:
        t = np.asarray(pending, dtype=np.uint64)
        x, y, _, _ = po.philox4x32_10(t, rnd, batch_id, (draw_index & 0xFFFFFFFF) ^ po.TAG_REPLAY, k0, k1)
        nxt = []
        for pos, a, b in zip(pending, x.tolist(), y.tolist()):
            v = (((a << 32) | b) * filled) >> 64
            if v in owner:              # claimed in an earlier round or by a smaller t of this round
                nxt.append(pos)
            else:
                owner[v] = pos
                out[pos] = v
        pending = nxt
        rnd += 1
    return out
