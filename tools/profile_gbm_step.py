"""One C4 step (state path) for ncu: python tools/profile_gbm_step.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rlmd_b200 import engine, lev_exp
n, h = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000, int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
lev = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), np.float32)
data_T = torch.empty((len(lev), n), dtype=torch.float32, device="cuda")
for it in range(2):
    res = engine.lev_sweep("gbm", lev, 100.0, n_investors=n, horizon=h, seed=420, log_mean=-0.05, sigma=0.2 ** 0.5,
                           mode="log", out_data_T=data_T, want_state=True)
    st = engine.rowstats(data_T, 1250)
    gs = engine.gbm_growth_summary(res["state"], lev, h, 100.0, data_T=data_T, quantiles=(0.05, 0.5))
torch.cuda.synchronize()
print(gs[:, 0].tolist())
