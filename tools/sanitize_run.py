"""
Small invocations of every kernel family for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Sizes are tiny (the tools slow kernels 10-100x); every family that uses shared memory, mbarriers,
cluster barriers or cross-block flags is touched: TMA chain kernels, packed / uint8 count kernels with the
tally sink, ingest, compaction, the cross-rank merge (all ranks played on this GPU), the cluster select,
row statistics, growth summaries, GBM Philox sweep + state summaries, env step, collector, replay.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as T

import __graft_entry__ as entry
from rlmd_b200 import _lib, collector, engine, envs, lev_exp, tally
from rlmd_b200._lib import check, lib, ptr, stream_ptr
from rlmd_b200.replay_torch import ReplayBufferTorch

entry.smoke()                                   # chain (TMA ring), log sweeps, packed, rowstats, series
rs = np.random.RandomState(1)
n, h = 3001, 203
oc = rs.choice(3, size=(n, h), p=[1 / 6, 1 / 6, 2 / 3])
lev = np.asarray(lev_exp.param_range(0.05, 1.0, 0.05), np.float32)
f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
for src in (T.tensor(oc.astype(np.int64)).cuda(), T.tensor(oc.astype(np.float32)), engine.encode_codes(oc.astype(np.uint8)),
            engine.pack_codes(engine.encode_codes(oc.astype(np.uint8)))):
    st = engine.lev_final_stats(f, 100.0, 3, src)            # ingest / count -> tally -> compaction -> cluster select
# the cross-rank merge with three ranks on this GPU (publish all, then merge all)
world, ranks = 3, []
for r in range(world):
    t = tally.FinalTally.__new__(tally.FinalTally)
    t.dev, t.group, t._staging, t.rows, t.horizon = T.device("cuda", 0), None, None, 0, None
    t.plan = _lib.TallyPlan()
    t.plan.rows_cap, t.plan.bins_cap, t.plan.grid_cap, t.plan.world = n, n, 64, world
    t.ws = T.empty((lib.b200_tally_workspace_bytes(C.byref(t.plan)) // 8,), dtype=T.int64, device="cuda")
    t.exchange = None
    t.ex_buf = T.zeros((lib.b200_tally_exchange_bytes(C.byref(t.plan)) // 8 + 1,), dtype=T.int64, device="cuda")
    check(lib.b200_tally_reset(C.byref(t.plan), ptr(t.ws), stream_ptr()))
    ranks.append(t)
cuts = [0, 1000, 1000, n]
peers = []
for r, t in enumerate(ranks):
    if cuts[r + 1] > cuts[r]:
        t.add(engine.encode_codes(oc[cuts[r]:cuts[r + 1]].astype(np.uint8)), 3)
    p = _lib.TallyPeers()
    p.world, p.rank, p.epoch = world, r, 1
    for q in range(world):
        p.exchange[q] = ranks[q].ex_buf.data_ptr()
    peers.append(p)
for ph in (0, 1):
    for r, t in enumerate(ranks):
        check(lib.b200_tally_finalize(C.byref(t.plan), ptr(t.ws), C.byref(peers[r]), ph, stream_ptr()))
got = [t.stats(f, 100.0, h, n_total=n, top=3) for t in ranks]
assert T.equal(got[0][:4], got[2][:4])
# GBM: Philox sweep, state summaries, generic growth, series
levg = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), np.float32)
res = engine.lev_sweep("gbm", levg, 100.0, n_investors=5000, horizon=130, seed=3, log_mean=-0.05, sigma=0.447, mode="log",
                       want_state=True)
engine.gbm_growth_summary(res["state"], levg, 130, 100.0, data_T=res["data_T"])
res = engine.lev_sweep("gbm", levg, 100.0, n_investors=5000, horizon=130, seed=3, log_mean=-0.05, sigma=0.447, mode="log",
                       want_log_w=True)
engine.growth_summary(res["log_w"], 130, 100.0, data_T=res["data_T"])
engine.rowstats(res["data_T"], 5)
x = engine.lev_draw("gbm", 2000, 96, seed=12, log_mean=-0.05, sigma=0.447)
engine.lev_series("gbm", levg, levg, 100.0, 2, outcomes=x, chunk_steps=32)
engine.lev_sweep("gbm", levg, 100.0, outcomes=x, mode="log")                 # TMA-staged fp32 tiles
# big brain, env step, collector, replay
codes = engine.encode_codes(rs.randint(0, 2, size=(1500, 40)).astype(np.uint8))
engine.bigbrain_series("coin", codes, 2, 100.0, (-0.4, 0.5), 2.5, [0.1, 0.5], [0.0, 0.7])
env = envs.Dice_SH_InvC(n_envs=777, seed=5)
a = T.full((777, env.action_dim), 0.3, dtype=T.float64, device="cuda")
for _ in range(3):
    env.step(a)
e1 = envs.Coin_InvA(1, n_envs=64, seed=1)
col = collector.Collector(e1, 200, {"mini_batch_size": 32, "discount": 0.99, "multi_steps": 5, "r_abs_zero": None,
                                    "dynamics": "M"}, seed=2)
act = T.full((64, 1), 0.4, dtype=T.float64, device="cuda")
for _ in range(4):
    col.step(act)
col.sample(2)
rb = ReplayBufferTorch({"gpu": "cuda:0", "input_dims": (5,), "num_actions": 1, "mini_batch_size": 64, "discount": 0.99,
                        "multi_steps": 5, "r_abs_zero": None, "dynamics": "M", "buffer": 4096, "n_cumsteps": 4096})
stt = T.randn((3000, 5), dtype=T.float64, device="cuda")
dn = T.zeros(3000, dtype=T.bool, device="cuda")
dn[::37] = True
rb.store_batch(stt, stt[:, :1], 1 + 0.01 * stt[:, 0], stt, dn)
rb.sample_exp()
rb.sample_many(8)
rb.capture_sampler()()
T.cuda.synchronize()
print("sanitize_run: all kernel families ran")
