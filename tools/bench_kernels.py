"""
Kernel-level timings on one B200 (CUDA events on the launching stream, L2
flushed by inputs much larger than L2).  Writes JSON lines to stdout.

    python tools/bench_kernels.py [--n 1000000] [--h 10000]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlmd_b200 import engine, lev_exp  # noqa: E402


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--h", type=int, default=10_000)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    n, h = a.n, a.h
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    hbm = peaks.get("hbm_gbs", 6650.0)

    def emit(name, best, med, steps, bytes_per_step, **kw):
        rec = dict(kernel=name, n=n, h=h, best_s=best, median_s=med, investor_steps_per_s=steps / best,
                   algorithmic_GBps=steps * bytes_per_step / best / 1e9,
                   hbm_frac=steps * bytes_per_step / best / 1e9 / hbm, **kw)
        print(json.dumps(rec), flush=True)

    dice = engine.lev_draw("discrete", n, h, seed=420, probs=(1 / 6, 1 / 6, 2 / 3))
    coin = None
    lev10 = np.asarray(lev_exp.param_range(0.1, 1.0, 0.1), np.float32)
    lev20 = np.asarray(lev_exp.param_range(0.05, 1.0, 0.05), np.float32)
    out10 = torch.empty((10, n), dtype=torch.float32, device="cuda")
    out20 = torch.empty((20, n), dtype=torch.float32, device="cuda")
    steps = n * h
    if a.only in ("", "chain"):
        for g, lev, out in ((10, lev10, out10), (20, lev20, out20)):
            f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
            for v in (1, 2, 3):
                b, m = timeit(lambda: engine.lev_sweep("discrete", f, 100.0, outcomes=dice, mode="chain",
                                                        variant=v, out_data_T=out))
                emit(f"chain_dice_G{g}_v{v}", b, m, steps, 1, G=g, path_steps_per_s=steps * g / b)
        coin = engine.lev_draw("discrete", n, h, seed=421, probs=(0.5, 0.5))
        f = lev_exp.coin_factor_table(lev10, 0.5, -0.4)
        for v in (1, 2, 3):
            b, m = timeit(lambda: engine.lev_sweep("discrete", f, 100.0, outcomes=coin, mode="chain", variant=v,
                                                    out_data_T=out10))
            emit(f"chain_coin_G10_v{v}", b, m, steps, 1, G=10, path_steps_per_s=steps * 10 / b)
        del coin
    if a.only in ("", "log"):
        f = lev_exp.dice_factor_table(lev10, 0.5, -0.5, 0.05)
        b, m = timeit(lambda: engine.lev_sweep("discrete", f, 100.0, outcomes=dice, mode="log", out_data_T=out10))
        emit("log_dice_G10", b, m, steps, 1, G=10)
    if a.only in ("", "grid2d"):
        # C3: die roll with safe-haven insurance, 2-D grid 32 x 32 (leverage x insurance fraction), packed outcomes:
        # one count pass + 16 wealth-from-counts launches, then the 12 statistics of the 1024 rows
        aa, bb = np.linspace(0.0, 1.2, 32, dtype=np.float32), np.linspace(0.0, 0.3, 32, dtype=np.float32)
        table = lev_exp.grid2d_factor_table(np.repeat(aa, 32), np.tile(bb, 32), (0.5, -0.5, 0.05), (-1.0, 5.0, -1.0))
        pk = engine.pack_codes(dice)
        b, m = timeit(lambda: engine.lev_grid_sweep(table, 100.0, pk), warm=1, reps=3)
        res = engine.lev_grid_sweep(table, 100.0, pk)
        b2, m2 = timeit(lambda: engine.rowstats(res["data_T"], max(1, n // 10000)), warm=1, reps=3)
        print(json.dumps(dict(kernel="grid2d_dice_sh_32x32_packed", n=n, h=h, grid_points=1024, sweep_s=b, stats_s=b2,
                              investor_steps_per_s=steps / (b + b2), path_steps_per_s=1024 * steps / (b + b2),
                              note="sweep = 1 count pass over 2.5 GB + 16 wealth-from-counts launches (4 GB of data_T); "
                                   "stats = 4 passes over the 1024 x 1e6 wealth rows")), flush=True)
        del pk, res
    if a.only in ("", "stats"):
        b, m = timeit(lambda: engine.rowstats(out10, max(1, n // 10000)))
        print(json.dumps(dict(kernel="rowstats_10rows", n=n, best_s=b, median_s=m,
                              GBps_4pass=4 * 10 * n * 4 / b / 1e9)), flush=True)
    if a.only in ("", "philox"):
        f = lev_exp.dice_factor_table(lev10, 0.5, -0.5, 0.05)
        for v in (1, 2, 3):
            b, m = timeit(lambda: engine.lev_sweep("discrete", f, 100.0, n_investors=n, horizon=h, seed=1,
                                                    probs=(1 / 6, 1 / 6, 2 / 3), mode="chain", variant=v,
                                                    out_data_T=out10), warm=1, reps=3)
            emit(f"chain_dice_philox_G10_v{v}", b, m, steps, 0, G=10, path_steps_per_s=steps * 10 / b)
    del dice
    if a.only in ("", "gbm"):
        levg = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), np.float32)
        ng = n // 4
        x = engine.lev_draw("gbm", ng, h, seed=3, log_mean=0.05 - 0.1, sigma=0.2 ** 0.5)
        outg = torch.empty((10, ng), dtype=torch.float32, device="cuda")
        b, m = timeit(lambda: engine.lev_sweep("gbm", levg, 100.0, outcomes=x, mode="log", out_data_T=outg))
        rec_steps = ng * h
        print(json.dumps(dict(kernel="gbm_stream_G10", n=ng, h=h, best_s=b, median_s=m,
                              investor_steps_per_s=rec_steps / b, algorithmic_GBps=rec_steps * 4 / b / 1e9,
                              hbm_frac=rec_steps * 4 / b / 1e9 / hbm)), flush=True)
        del x
        npx = 12_500_000 if n >= 1_000_000 else n
        outp = torch.empty((10, npx), dtype=torch.float32, device="cuda")
        b, m = timeit(lambda: engine.lev_sweep("gbm", levg, 100.0, n_investors=npx, horizon=h, seed=3,
                                                log_mean=0.05 - 0.1, sigma=0.2 ** 0.5, mode="log",
                                                out_data_T=outp), warm=1, reps=3)
        print(json.dumps(dict(kernel="gbm_philox_G10", n=npx, h=h, best_s=b, median_s=m,
                              investor_steps_per_s=npx * h / b)), flush=True)

    if a.only in ("", "series"):
        # *_smart_lev: chain + per-step dump + 12 statistics per (leverage, step); H reduced to keep it short
        f = lev_exp.dice_factor_table(lev10, 0.5, -0.5, 0.05)
        for hs in (min(h, 512), min(h, 4096)):
            oc = engine.lev_draw("discrete", n, hs, seed=420, probs=(1 / 6, 1 / 6, 2 / 3))
            b, m = timeit(lambda: engine.lev_series("discrete", f, lev10, 100.0, max(1, n // 10000), outcomes=oc),
                          warm=1, reps=3)
            ps = n * hs * 10
            print(json.dumps(dict(kernel="series_dice_G10", n=n, h=hs, best_s=b, median_s=m,
                                  investor_steps_per_s=n * hs / b, path_steps_per_s=ps / b,
                                  bytes_per_path_step=20, GBps_dump_plus_4pass=ps * 20 / b / 1e9,
                                  hbm_frac=ps * 20 / b / 1e9 / hbm)), flush=True)
            del oc
    if a.only in ("", "bigbrain"):
        hs = min(h, 256)
        oc = engine.lev_draw("discrete", n, hs, seed=5, probs=(0.5, 0.5))
        stop = np.asarray(lev_exp.param_range(0.05, 0.95, 0.05), np.float32)
        roll = np.asarray(lev_exp.param_range(0.70, 0.95, 0.05), np.float32)
        for kind, rets in (("coin", (-0.4, 0.5)), ("dice", (0.5, -0.5, 0.05))):
            b, m = timeit(lambda: engine.bigbrain_series(kind, oc, max(1, n // 10000), 100.0, rets, 2.5, stop[:4],
                                                         roll[:2]), warm=1, reps=2)
            ps = n * hs * 8
            print(json.dumps(dict(kernel=f"bigbrain_{kind}_P8", n=n, h=hs, best_s=b, median_s=m,
                                  path_steps_per_s=ps / b, bytes_per_path_step=40,
                                  hbm_frac=ps * 40 / b / 1e9 / hbm)), flush=True)
        del oc
    if a.only in ("", "env"):
        from rlmd_b200 import envs
        for e in (1, 100, 100_000, 10_000_000):
            env = envs.Coin_InvA(1, n_envs=e) if e > 1 else envs.Coin_InvA(1)
            env.reset()
            if e > 1:
                act = torch.rand((e, 1), dtype=torch.float64, device="cuda") * 1.98 - 0.99
                fn = lambda: env.step(act)
            else:
                fn = lambda: env.step(np.array([0.3]))
            b, m = timeit(fn, warm=3, reps=20)
            print(json.dumps(dict(kernel="menv_step_coin_A1", n_envs=e, best_s=b, median_s=m, env_steps_per_s=e / b,
                                  algorithmic_GBps=e * 130 / b / 1e9, hbm_frac=e * 130 / b / 1e9 / hbm)), flush=True)
    if a.only in ("", "replay"):
        from rlmd_b200.replay_torch import ReplayBufferTorch
        mem = 1_000_000
        rs = np.random.RandomState(0)
        lens = rs.randint(5, 61, size=mem // 5)
        done = np.zeros(mem, dtype=bool)
        done[np.cumsum(lens)[np.cumsum(lens) < mem] - 1] = True
        for nstep in (1, 5, 10):
            for batch in (256, 512):
                inputs = {"gpu": "cuda:0", "input_dims": (5,), "num_actions": 1, "mini_batch_size": batch,
                          "discount": 0.99, "multi_steps": nstep, "r_abs_zero": None, "dynamics": "M", "buffer": mem,
                          "n_cumsteps": mem}
                buf = ReplayBufferTorch(inputs)
                st = torch.randn((mem, 5), dtype=torch.float64, device="cuda")
                t0 = timeit(lambda: None, warm=0, reps=1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                buf.store_batch(st, st[:, :1], 1 + 0.01 * st[:, 0], st, torch.as_tensor(done, device="cuda"))
                e1.record()
                e1.synchronize()
                store_s = e0.elapsed_time(e1) * 1e-3
                b1, m1 = timeit(lambda: buf.sample_exp(), warm=5, reps=50)
                k = 1024
                bk, mk = timeit(lambda: buf.sample_many(k), warm=2, reps=10)
                k2 = 16384        # 4.2e6 (batch 256) samples per launch: the kernels, not the Python call, set the time
                bk2, _ = timeit(lambda: buf.sample_many(k2), warm=2, reps=10)
                pre = torch.randint(0, mem, (k2, batch), device="cuda")
                bg, _ = timeit(lambda: buf.sample_many(k2, batches=pre), warm=2, reps=10)
                bytes_per = 106 if nstep == 1 else 114 + 4 * (nstep - 1)
                print(json.dumps(dict(kernel="replay", n_step=nstep, batch=batch, buffer=mem, store_1e6_s=store_s,
                                      stores_per_s=mem / store_s, sample_call_us=b1 * 1e6, samples_per_s_one_call=batch / b1,
                                      many_batches=k, samples_per_s_many=k * batch / bk,
                                      algorithmic_GBps_many=k * batch * bytes_per / bk / 1e9,
                                      hbm_frac_many=k * batch * bytes_per / bk / 1e9 / hbm,
                                      samples_per_s_16384_batches=k2 * batch / bk2,
                                      gather_only_samples_per_s=k2 * batch / bg,
                                      gather_only_algorithmic_GBps=k2 * batch * bytes_per / bg / 1e9,
                                      gather_only_hbm_frac=k2 * batch * bytes_per / bg / 1e9 / hbm)), flush=True)
                del buf, st
    if a.only in ("", "collect"):
        # fused collector: env step + append per environment (one kernel), then an n-step sample of 256;
        # eager = one C call per stage from Python, graph = the captured pair replayed
        from rlmd_b200 import collector, envs
        for e in (1, 256, 65_536, 1_048_576):
            lane = 1_000_000 if e == 1 else max(64, 64_000_000 // e)
            env = envs.Coin_InvA(1, n_envs=e, seed=1)
            inputs = {"mini_batch_size": 256, "discount": 0.99, "multi_steps": 5, "r_abs_zero": None, "dynamics": "M"}
            col = collector.Collector(env, lane, inputs, seed=2)
            act = torch.rand((e, 1), dtype=torch.float64, device="cuda") * 0.8 + 0.1
            for _ in range(8):
                col.step(act)
            reps = min(200, lane - 20)
            def eager():
                col.step(act)
                col.sample(1)
            be, me = timeit(eager, warm=3, reps=reps // 4)
            run = col.capture(act, k=1)
            bg, mg = timeit(run, warm=3, reps=reps // 4)
            def steps_only():
                col.step(act)
            bs, ms = timeit(steps_only, warm=3, reps=reps // 4)
            bytes_step = e * (8 + 8 * 5 * 2 + 8 + 53 + 8 * 5 + 8 + 2 + 32)
            print(json.dumps(dict(kernel="collect_step_sample", n_envs=e, lane_len=lane, eager_us=me * 1e6,
                                  graph_us=mg * 1e6, step_only_us=ms * 1e6, env_steps_per_s_graph=e / mg,
                                  env_steps_per_s_eager=e / me, env_steps_per_s_step_only=e / ms,
                                  step_GBps=bytes_step / ms / 1e9)), flush=True)
            del col, env
        env = envs.Coin_InvA(1, seed=3)
        for n_eval, steps in ((100, 1000), (100_000, 1000), (4_000_000, 200)):
            a_ev = torch.full((n_eval, 1), 0.25, dtype=torch.float64, device="cuda")
            b, m = timeit(lambda: collector.rollout(env, a_ev, steps), warm=1, reps=5)
            _, st, _, _ = collector.rollout(env, a_ev, steps)
            tot = float(st.sum())
            print(json.dumps(dict(kernel="eval_rollout_coin_A1", n_eval=n_eval, max_steps=steps, best_s=b,
                                  env_steps_per_s=tot / b, mean_steps=tot / n_eval)), flush=True)
    if a.only in ("", "growth"):
        lw = torch.randn((20, n), dtype=torch.float64, device="cuda")
        b, m = timeit(lambda: engine.growth_summary(lw, h, 100.0, quantiles=(0.05, 0.5, 0.95)))
        print(json.dumps(dict(kernel="growth_summary_20rows", n=n, best_s=b, median_s=m,
                              GBps_6pass=6 * 20 * n * 8 / b / 1e9)), flush=True)


if __name__ == "__main__":
    main()
