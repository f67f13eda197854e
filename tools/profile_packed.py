"""
The bench's sweep launch at full size (1e6 investors x 1e4 steps, 20 leverages) on packed
and uint8 outcomes - the program `ncu --set full -k regex:log_discrete` wraps for
roofline.traffic (profiles/roofline_traffic.json).  Nothing is timed here.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlmd_b200 import engine, lev_exp  # noqa: E402

n, h = 1_000_000, 10_000
lev = np.asarray(lev_exp.param_range(0.05, 1.00, 0.05), np.float32)
f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
out = torch.empty((len(lev), n), dtype=torch.float32, device="cuda")
pk = engine.lev_draw("discrete", n, h, seed=420, probs=(1 / 6, 1 / 6, 2 / 3), packed=True)
for _ in range(2):
    engine.lev_sweep("discrete", f, 100.0, outcomes=pk, mode="log", out_data_T=out)
if "--u8" in sys.argv:
    u8 = engine.lev_draw("discrete", n, h, seed=420, probs=(1 / 6, 1 / 6, 2 / 3))
    engine.lev_sweep("discrete", f, 100.0, outcomes=u8, mode="log", out_data_T=out)
torch.cuda.synchronize()
print("profile program ok")
