"""
One launch of every kernel family at a moderate size - the program ncu wraps
(`ncu --set full -k regex:... python tools/profile_kernels.py`).  Nothing here
is timed; numbers printed under ncu are never bench values.

    python tools/profile_kernels.py [--only chain,gbm,...]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlmd_b200 import engine, lev_exp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--h", type=int, default=2048)
    ap.add_argument("--series-h", type=int, default=64,
                    help="steps of the series run (chunks of 32 steps; early chunks are tie-heavy: with a long run, "
                         "point ncu at late chunks with --launch-skip)")
    a = ap.parse_args()
    only = set(filter(None, a.only.split(",")))
    n, h = a.n, a.h

    def on(name):
        return not only or name in only

    lev10 = np.asarray(lev_exp.param_range(0.1, 1.0, 0.1), np.float32)
    levg = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), np.float32)
    if on("chain") or on("log") or on("stats"):
        dice = engine.lev_draw("discrete", n, h, seed=420, probs=(1 / 6, 1 / 6, 2 / 3))
        f3 = lev_exp.dice_factor_table(lev10, 0.5, -0.5, 0.05)
        out = torch.empty((10, n), dtype=torch.float32, device="cuda")
        if on("chain"):
            for v in (1, 2, 3):
                try:
                    engine.lev_sweep("discrete", f3, 100.0, outcomes=dice, mode="chain", variant=v, out_data_T=out)
                except Exception as e:  # a variant this build does not have
                    print("chain variant", v, "skipped:", e)
            coin = engine.lev_draw("discrete", n, h, seed=421, probs=(0.5, 0.5))
            f2 = lev_exp.coin_factor_table(lev10, 0.5, -0.4)
            for v in (1, 3):
                try:
                    engine.lev_sweep("discrete", f2, 100.0, outcomes=coin, mode="chain", variant=v, out_data_T=out)
                except Exception as e:
                    print("chain variant", v, "skipped:", e)
            del coin
        if on("log"):
            engine.lev_sweep("discrete", f3, 100.0, outcomes=dice, mode="log", out_data_T=out)
        if on("stats"):
            engine.lev_sweep("discrete", f3, 100.0, outcomes=dice, mode="log", out_data_T=out)
            engine.rowstats(out, max(1, n // 10000))
        del dice
    if on("gbm"):
        x = engine.lev_draw("gbm", n // 4, h, seed=3, log_mean=-0.05, sigma=0.2 ** 0.5)
        outg = torch.empty((10, n // 4), dtype=torch.float32, device="cuda")
        engine.lev_sweep("gbm", levg, 100.0, outcomes=x, mode="log", out_data_T=outg)
        del x
        outp = torch.empty((10, n), dtype=torch.float32, device="cuda")
        engine.lev_sweep("gbm", levg, 100.0, n_investors=n, horizon=h, seed=3, log_mean=-0.05, sigma=0.2 ** 0.5,
                         mode="log", out_data_T=outp)
    if on("series"):
        f3 = lev_exp.dice_factor_table(lev10, 0.5, -0.5, 0.05)
        oc = engine.lev_draw("discrete", n, a.series_h, seed=420, probs=(1 / 6, 1 / 6, 2 / 3))
        engine.lev_series("discrete", f3, lev10, 100.0, max(1, n // 10000), outcomes=oc)
        del oc
    if on("bigbrain"):
        oc = engine.lev_draw("discrete", n, 32, seed=5, probs=(0.5, 0.5))
        stop = np.asarray(lev_exp.param_range(0.05, 0.95, 0.05), np.float32)
        roll = np.asarray(lev_exp.param_range(0.70, 0.95, 0.05), np.float32)
        engine.bigbrain_series("coin", oc, max(1, n // 10000), 100.0, (-0.4, 0.5), 2.5, stop[:4], roll[:2])
        del oc
    if on("env"):
        from rlmd_b200 import envs
        e = 4_000_000
        env = envs.Coin_InvA(1, n_envs=e)
        env.reset()
        act = torch.rand((e, 1), dtype=torch.float64, device="cuda") * 1.98 - 0.99
        env.step(act)
    if on("replay"):
        from rlmd_b200.replay_torch import ReplayBufferTorch
        mem = 1_000_000
        rs = np.random.RandomState(0)
        lens = rs.randint(5, 61, size=mem // 5)
        done = np.zeros(mem, dtype=bool)
        ends = np.cumsum(lens)
        done[ends[ends < mem] - 1] = True
        for nstep in (1, 10):
            inputs = {"gpu": "cuda:0", "input_dims": (5,), "num_actions": 1, "mini_batch_size": 256,
                      "discount": 0.99, "multi_steps": nstep, "r_abs_zero": None, "dynamics": "M", "buffer": mem,
                      "n_cumsteps": mem}
            buf = ReplayBufferTorch(inputs)
            st = torch.randn((mem, 5), dtype=torch.float64, device="cuda")
            buf.store_batch(st, st[:, :1], 1 + 0.01 * st[:, 0], st, torch.as_tensor(done, device="cuda"))
            buf.sample_exp()
            idx = buf.sample_many(16384)[0]                  # one launch: draw + gather
            buf.sample_many(16384, batches=idx)              # the warp-cooperative gather alone
    if on("collect"):
        from rlmd_b200 import collector, envs
        e = 1_048_576
        env = envs.Coin_InvA(1, n_envs=e, seed=1)
        inputs = {"mini_batch_size": 256, "discount": 0.99, "multi_steps": 5, "r_abs_zero": None, "dynamics": "M"}
        col = collector.Collector(env, 64, inputs, seed=2)
        act = torch.rand((e, 1), dtype=torch.float64, device="cuda") * 0.8 + 0.1
        for _ in range(3):
            col.step(act)
        col.sample(1024)
        collector.rollout(envs.Coin_InvA(1, seed=3), torch.full((1_000_000, 1), 0.25, dtype=torch.float64,
                                                                 device="cuda"), 200)
    if on("growth"):
        lw = torch.randn((20, n), dtype=torch.float64, device="cuda")
        engine.growth_summary(lw, h, 100.0, quantiles=(0.05, 0.5, 0.95))
    torch.cuda.synchronize()
    print("profile program ok")


if __name__ == "__main__":
    main()
