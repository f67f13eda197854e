"""The multi-GPU tally finalize played on ONE GPU (all ranks publish, then all merge) at the bench's size,
so that ncu can time the merge kernels: python tools/profile_tally_merge.py [world]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as T

from rlmd_b200 import _lib, engine, lev_exp, tally
from rlmd_b200._lib import check, lib, ptr, stream_ptr

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, h = 1_000_000, 10_000
lev = np.asarray(lev_exp.param_range(0.05, 1.0, 0.05), np.float32)
f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
ranks, ocs = [], []
for r in range(world):
    t = tally.FinalTally.__new__(tally.FinalTally)
    t.dev, t.group, t._staging, t.rows, t.horizon = T.device("cuda", 0), None, None, 0, None
    t.plan = _lib.TallyPlan()
    t.plan.rows_cap, t.plan.bins_cap, t.plan.grid_cap, t.plan.world = n, min(n * world, tally.DEFAULT_BINS_CAP), 64, world
    t.ws = T.empty((lib.b200_tally_workspace_bytes(C.byref(t.plan)) // 8,), dtype=T.int64, device="cuda")
    t.exchange = None
    t.ex_buf = T.zeros((lib.b200_tally_exchange_bytes(C.byref(t.plan)) // 8 + 1,), dtype=T.int64, device="cuda")
    check(lib.b200_tally_reset(C.byref(t.plan), ptr(t.ws), stream_ptr()))
    ranks.append(t)
    ocs.append(engine.lev_draw("discrete", n, h, seed=420, investor_offset=r * n, probs=(1 / 6, 1 / 6, 2 / 3), packed=True))
for epoch in (1, 2, 3):
    peers = []
    for r, t in enumerate(ranks):
        t.rows = 0
        t.add(ocs[r], 3)
        p = _lib.TallyPeers()
        p.world, p.rank, p.epoch = world, r, epoch
        for q in range(world):
            p.exchange[q] = ranks[q].ex_buf.data_ptr()
        peers.append(p)
    for r, t in enumerate(ranks):
        check(lib.b200_tally_finalize(C.byref(t.plan), ptr(t.ws), C.byref(peers[r]), 0, stream_ptr()))
    for r, t in enumerate(ranks):
        check(lib.b200_tally_finalize(C.byref(t.plan), ptr(t.ws), C.byref(peers[r]), 1, stream_ptr()))
    st = [t.stats(f, 100.0, h, n_total=n * world, top=100 * world) for t in ranks]
T.cuda.synchronize()
print(ranks[0].info(), bool(T.equal(st[0], st[-1])))
