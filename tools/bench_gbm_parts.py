"""Parts of the C4 step (GBM Philox sweep, row statistics, growth summaries) timed separately with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rlmd_b200 import engine, lev_exp
n, h = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000, 10_000
lev = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), np.float32)
data_T = torch.empty((len(lev), n), dtype=torch.float32, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(3):
    e = [ev() for _ in range(4)]
    e[0].record()
    res = engine.lev_sweep("gbm", lev, 100.0, n_investors=n, horizon=h, seed=420, log_mean=-0.05, sigma=0.2 ** 0.5,
                           mode="log", out_data_T=data_T, want_log_w=True)
    e[1].record()
    st = engine.rowstats(data_T, 1250)
    e[2].record()
    gs = engine.growth_summary(res["log_w"], h, 100.0, data_T=data_T, quantiles=(0.05, 0.5))
    e[3].record()
    torch.cuda.synchronize()
    print("sweep / rowstats / growth ms:", [round(e[i].elapsed_time(e[i + 1]), 3) for i in range(3)])

# the state path of the bench step: sweep (state out) -> rowstats -> gbm_growth_summary
from rlmd_b200 import _lib
for it in range(3):
    e = [ev() for _ in range(4)]
    e[0].record()
    res = engine.lev_sweep("gbm", lev, 100.0, n_investors=n, horizon=h, seed=420, log_mean=-0.05, sigma=0.2 ** 0.5,
                           mode="log", out_data_T=data_T, want_state=True)
    e[1].record()
    st = engine.rowstats(data_T, 1250)
    e[2].record()
    gs = engine.gbm_growth_summary(res["state"], lev, h, 100.0, data_T=data_T, quantiles=(0.05, 0.5))
    e[3].record()
    torch.cuda.synchronize()
    print("state path: sweep / rowstats / gbm_growth ms:", [round(e[i].elapsed_time(e[i + 1]), 3) for i in range(3)])
