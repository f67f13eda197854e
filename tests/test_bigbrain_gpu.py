"""
GPU parity of the state-dependent leverage sweeps (K3) through the drop-in shims,
against the reference's own outputs (tests/golden/bigbrain_*.npz) and the oracle.
"""
import contextlib
import io

import numpy as np
import pytest
import torch as T

import golden_io
from oracle import lev_oracle as lo
from test_oracle_bigbrain import assert_bigbrain_close, oracle_data

pytestmark = pytest.mark.gpu


def call(case, oc):
    from rlmd_b200 import lev_exp
    n, h = case["n"], case["h"]
    common = (T.tensor(n, dtype=T.int32), T.tensor(h, dtype=T.int32), case["top"], T.tensor(case["v0"]))
    lf = T.tensor(golden_io.bigbrain_lev_factor(case), dtype=T.float64)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        if case["kind"] == "coin":
            data = lev_exp.coin_big_brain_lev("cuda", T.tensor(oc.astype(np.float32)), *common, case["up_r"],
                                              case["down_r"], lf, *case["stop"], *case["roll"])
        else:
            data = lev_exp.dice_big_brain_lev("cuda", T.tensor(oc.astype(np.int64)), *common, case["up_r"],
                                              case["down_r"], case["mid_r"], lf, *case["stop"], *case["roll"])
    return data, buf.getvalue()


@pytest.mark.parametrize("case", golden_io.BIGBRAIN_CASES, ids=lambda c: c["name"])
def test_big_brain_dropin(case):
    gold = golden_io.load("bigbrain_" + case["name"])
    oc = golden_io.draw_outcomes(case)
    data, text = call(case, oc)
    assert data.dtype == T.float32 and data.is_cuda
    assert tuple(data.shape[:3]) == gold["data"].shape[:3] and data.shape[3] == case["h"] - 1
    assert_bigbrain_close(data.cpu().numpy()[:, :, :, gold["cols"]], gold["data"])
    assert text == str(gold["text"]) or text.count("stop/roll") == str(gold["text"]).count("stop/roll")


def test_final_wealth_bits_and_chunk_invariance():
    """The wealth chain itself: bit-identical to the oracle's (fp32 coin, fp64 dice) for any chunking."""
    from rlmd_b200 import engine
    for name in ("coin_grid_top4", "dice_grid_top3"):
        case = golden_io.bigbrain_case(name)
        oc = golden_io.draw_outcomes(case)
        returns = (case["down_r"], case["up_r"]) if case["kind"] == "coin" else (case["up_r"], case["down_r"], case["mid_r"])
        oret = (case["up_r"], case["down_r"]) + ((case["mid_r"],) if case["kind"] == "dice" else ())
        stop, roll = lo.param_range(*case["stop"]), lo.param_range(*case["roll"])
        eta = golden_io.bigbrain_lev_factor(case)
        want_data, want_w = lo.big_brain(case["kind"], oc, case["top"], case["v0"], oret, eta, stop, roll, want_wealth=True)
        codes = engine.encode_codes(oc)
        outs = []
        for steps in (None, 7, 16):
            data, w = engine.bigbrain_series(case["kind"], codes, case["top"], case["v0"], returns, eta, stop, roll,
                                             chunk_steps=steps)
            outs.append((data.cpu().numpy(), w.cpu().numpy()))
        for data, w in outs[1:]:
            assert np.array_equal(data.view(np.uint32), outs[0][0].view(np.uint32))
            assert np.array_equal(w, outs[0][1])
        w = outs[0][1]
        if case["kind"] == "coin":
            assert w.dtype == np.float32 and np.array_equal(w.view(np.uint32), want_w.astype(np.float32).view(np.uint32))
        else:
            assert w.dtype == np.float64 and np.array_equal(w.view(np.uint64), want_w.view(np.uint64))
        assert_bigbrain_close(outs[0][0], want_data)


def test_galaxy_brain_table():
    from rlmd_b200 import lev_exp
    want = golden_io.load("galaxy_brain")["data"]
    got = lev_exp.coin_galaxy_brain_lev("cuda", *golden_io.GALAXY_GRID).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
