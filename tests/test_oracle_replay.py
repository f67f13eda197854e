"""
The replay oracle (oracle/replay_oracle.py) against fixtures produced by the
unmodified reference ReplayBufferTorch (tests/golden/gen_golden_more.py:gen_replay),
and the closed form of SURVEY.md App. D against the oracle's list bookkeeping.
"""
import numpy as np
import pytest

import golden_io
from oracle import replay_oracle as ro


def _run(case, on_event):
    st = golden_io.replay_stream(case)
    buf = ro.ReplayOracle(golden_io.replay_inputs_dict(case))
    ev = 0
    for i in range(case["fill"] + 1):
        while ev < len(case["events"]) and case["events"][ev] == i:
            on_event(ev, buf, st["batches"][ev])
            ev += 1
        if i == case["fill"]:
            break
        buf.store_exp(st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i],
                      bool(st["done"][i]))
    return buf


@pytest.mark.parametrize("case", golden_io.REPLAY_CASES, ids=lambda c: c["name"])
def test_oracle_matches_reference_fixture(case):
    gold = golden_io.load("replay_" + case["name"])

    def check(ev, buf, batch):
        s, a, r, s2, d, eff = buf.sample_exp(batch)
        assert np.array_equal(batch, gold[f"ev{ev}_batch"])
        assert np.array_equal(eff, gold[f"ev{ev}_eff"]), f"event {ev}: effective lengths"
        assert np.array_equal(s.view(np.uint32), gold[f"ev{ev}_states"].view(np.uint32)), f"event {ev}: states"
        assert np.array_equal(a.view(np.uint32), gold[f"ev{ev}_actions"].view(np.uint32)), f"event {ev}: actions"
        assert np.array_equal(s2.view(np.uint32), gold[f"ev{ev}_next_states"].view(np.uint32))
        assert np.array_equal(d, gold[f"ev{ev}_dones"])
        # torch.sum / torch.prod over <= 9 fp32 terms: order is torch's; 1e-6 relative
        np.testing.assert_allclose(r, gold[f"ev{ev}_rewards"], rtol=2e-6, atol=1e-7)

    buf = _run(case, check)
    assert buf.mem_idx == int(gold["mem_idx"])
    assert np.array_equal(buf.reward_memory[: case["fill"]].view(np.uint32), gold["reward_memory"].view(np.uint32))


@pytest.mark.parametrize("case", [c for c in golden_io.REPLAY_CASES if c["n"] > 1], ids=lambda c: c["name"])
def test_closed_form_equals_list_bookkeeping(case):
    def check(ev, buf, batch):
        for step in range(buf.mem_idx):
            assert ro.closed_form_history(step, buf.terminal_memory, buf.mem_idx) == buf._construct_history(step), \
                (ev, step)

    _run(case, check)


def test_toy_buffer_of_the_survey():
    """SURVEY.md App. D toy check: lens 4,3,5 + 2 running, r_k = k+1, gamma .5, n = 3, additive."""
    inputs = {"gpu": "cpu", "input_dims": (1,), "num_actions": 1, "mini_batch_size": 14, "discount": 0.5,
              "multi_steps": 3, "r_abs_zero": None, "dynamics": "A", "buffer": 20, "n_cumsteps": 20}
    buf = ro.ReplayOracle(inputs)
    ends = {3, 6, 11}
    for k in range(14):
        buf.store_exp([k], [k], float(k + 1), [k], k in ends)
    _, _, r, _, _, eff = buf.sample_exp(np.arange(14))
    assert list(r) == [0, 1, 2, 3.5, 5, 8, 8, 8, 12.5, 14, 15.5, 15.5, 1, 2]
    assert list(eff) == [1, 2, 3, 3, 2, 3, 3, 2, 3, 3, 3, 3, 2, 3]
