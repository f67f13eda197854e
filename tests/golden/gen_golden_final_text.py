"""
Writes tests/golden/final_text.json: the text the UNMODIFIED reference's
`{coin,dice,dice_sh,gbm}_fixed_final_lev` (lev/lev_exp.py:56-125, :508-583,
:935-1005, :1121-1206) prints on the seeded outcomes of tests/golden_io.LEV_CASES,
called the way lev/*.py call them (CPU device, reference dtypes).

Run in the build container only (the reference tree does not travel):

    python tests/golden/gen_golden_final_text.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
import golden_io  # noqa: E402


def main():
    import torch as T

    lx = ref_shim.load("lev.lev_exp")
    dev = T.device("cpu")
    out = {}
    for case in golden_io.LEV_CASES:
        oc = golden_io.draw_outcomes(case)
        v0, top, grid = T.tensor(case["v0"]), case["top"], case["grid"]
        with ref_shim.quiet() as buf:
            if case["kind"] == "coin":
                lx.coin_fixed_final_lev(dev, T.tensor(oc.astype(np.float32)), top, v0, case["up_r"], case["down_r"],
                                        *grid)
            elif case["kind"] == "dice":
                lx.dice_fixed_final_lev(dev, T.tensor(oc.astype(np.int64)), top, v0, case["up_r"], case["down_r"],
                                        case["mid_r"], *grid)
            elif case["kind"] == "dice_sh":
                lx.dice_sh_fixed_final_lev(dev, T.tensor(oc.astype(np.int64)), top, v0, case["up_r"], case["down_r"],
                                           case["mid_r"], *case["sh"], *grid)
            else:
                lx.gbm_fixed_final_lev(dev, T.tensor(oc), top, v0, *grid)
        out[case["name"]] = buf.getvalue().rstrip("\n")
        print(case["name"], len(out[case["name"]].splitlines()), "lines")
    with open(os.path.join(HERE, "final_text.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
