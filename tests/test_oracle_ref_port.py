"""The torch-CPU baseline port prints exactly what the unmodified reference prints."""
import numpy as np
import pytest
import torch as T

import golden_io
from oracle import lev_ref_port as port
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.mark.parametrize("name", ["coin_top7", "dice_top5", "dicesh_top3", "gbm_snp_top4"])
def test_port_prints_reference_text(name):
    case = golden_io.lev_case(name)
    oc = golden_io.draw_outcomes(case)
    lx = ref_shim.load("lev.lev_exp")
    dev, v0, top, grid = T.device("cpu"), T.tensor(case["v0"]), case["top"], case["grid"]
    with ref_shim.quiet() as buf:
        if case["kind"] == "coin":
            t = T.tensor(oc.astype(np.float32))
            lx.coin_fixed_final_lev(dev, t, top, v0, case["up_r"], case["down_r"], *grid)
            rows, levs = port.fixed_final("coin", t, top, v0, (case["up_r"], case["down_r"]), grid)
        elif case["kind"] == "dice":
            t = T.tensor(oc.astype(np.int64))
            lx.dice_fixed_final_lev(dev, t, top, v0, case["up_r"], case["down_r"], case["mid_r"], *grid)
            rows, levs = port.fixed_final("dice", t, top, v0, (case["up_r"], case["down_r"], case["mid_r"]), grid)
        elif case["kind"] == "dice_sh":
            t = T.tensor(oc.astype(np.int64))
            lx.dice_sh_fixed_final_lev(dev, t, top, v0, case["up_r"], case["down_r"], case["mid_r"], *case["sh"], *grid)
            rows, levs = port.fixed_final("dice_sh", t, top, v0, (case["up_r"], case["down_r"], case["mid_r"]), grid,
                                          sh=case["sh"])
        else:
            t = T.tensor(oc)
            lx.gbm_fixed_final_lev(dev, t, top, v0, *grid)
            rows, levs = port.fixed_final("gbm", t, top, v0, None, grid)
    assert buf.getvalue().rstrip("\n") == port.format_rows(rows, levs)
