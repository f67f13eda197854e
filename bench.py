#!/usr/bin/env python
"""
bench.py - headline benchmark of the leverage-sweep hot path (contract in the
task prompt; BASELINE.json metric "investor-steps/sec (leverage sweep)").

Workload (BASELINE.json configs[1]): lev/dice_roll.py's trinary die-roll sweep,
1e6 investors x 1e4 steps per GPU, the script's final-time grid
param_range(0.05, 1.00, 0.05) = 20 leverages, top = 100.  One "step" is one
pass of the `dice_fixed_final_lev` hot path over one synthetic outcome array:
    sweep (log-domain count kernel) -> data_T[20, N] -> 12 summary statistics
    per leverage (exact order statistics by radix select).
The outcome array is resident in HBM in the engine's packed format (2 bits per
roll, `--format packed2`, the default; `--format u8` = one byte per roll): the
sweep is bound by the one read of that array, so its size is the cost.
Steps run strictly one after the other (sweep, then its statistics); the rate of
the same steps through engine.FinalSweepPipeline (statistics of step i beside
the sweep of step i+1) is reported as `pipelined` (`--pipeline` times that mode).
With N GPUs (one process per GPU, torchrun) every rank owns its own 1e6
investors (weak scaling); the only cross-GPU traffic is the all-reduce of the
per-leverage partial sums and radix histograms inside the statistics.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line.  `--impl reference` times the reference's own CPU
implementation of the same path (torch-CPU port, all host threads) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_INVESTORS = 1_000_000
HORIZON = 10_000
TOP = 100
GRID = (0.05, 1.00, 0.05)
RETURNS = (0.5, -0.5, 0.05)
PROBS = (1 / 6, 1 / 6, 2 / 3)
V0 = 100.0
METRIC = "investor-steps/sec (leverage sweep)"
UNIT = "investor-steps/s"
WORKLOAD = "lev/dice_roll.py dice_fixed_final_lev: 1e6 investors x 1e4 steps per GPU, 20 leverages, top 100"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------ CPU arm
def cpu_reference_step(n_s, h, threads=None):
    """One pass of the torch-CPU port over a fresh [n_s, h] int64 outcome array."""
    import torch

    from oracle import lev_ref_port as port

    if threads:
        torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(420)
    u = torch.rand((n_s, h), generator=gen)
    outcomes = torch.where(u < PROBS[0], 0, torch.where(u < PROBS[0] + PROBS[1], 1, 2)).to(torch.int64)
    del u
    top = max(1, int(n_s * 1e-4))
    t0 = time.perf_counter()
    rows, levs = port.fixed_final("dice", outcomes, top, V0, RETURNS, GRID)
    dt = time.perf_counter() - t0
    assert len(rows) == 20
    return dt


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = torch.get_num_threads()
    h = HORIZON
    # size the per-step sample so that the whole K+W run ends within ~2.5 minutes
    probe = cpu_reference_step(200, h)
    rate = 200 / probe  # investors per second at this horizon and grid
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_s = int(max(200, min(args.ref_sample, rate * budget)))
    for _ in range(args.warmup):
        cpu_reference_step(n_s, h)
    ts = [cpu_reference_step(n_s, h) for _ in range(args.steps)]
    total = sum(ts)
    value = n_s * h * args.steps / total
    sample = f"{n_s} investors x {h} steps per step (of the 1e6 x 1e4 workload), 20 leverages, torch-CPU ops"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------- clock sampler
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {
                0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown",
            }
            for bit, name in names.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._halt.is_set():
            self.sample()
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=1)
        import statistics

        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------- the metric's other clauses
def _event_time(torch, fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


def run_secondary(torch, dist, engine, lev_exp, np, dev, rank, world):
    """
    Side numbers for the other clauses of BASELINE.json's metric, outside the timed
    region of the headline (every rank runs them; times are the max over ranks):
    * C4: GBM leverage sweep, on-device Philox draws, 1.25e7 investors x 1e4 steps per
      GPU (1e8 investors over 8 GPUs), grid param_range(-1, 1, 0.2);
    * C5: n-step replay sampling, batch 256 from a full 1e6 buffer (per call, and 1024
      mini-batches per launch), and the fused collector (env step + append + sample).
    """
    from rlmd_b200 import collector, envs
    from rlmd_b200.replay_torch import ReplayBufferTorch

    out = {}
    n_g, h = 12_500_000, HORIZON
    levg = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), np.float32)
    buf = torch.empty((len(levg), n_g), dtype=torch.float32, device=dev)
    t = _event_time(torch, lambda: engine.lev_sweep(
        "gbm", levg, V0, n_investors=n_g, horizon=h, seed=420, investor_offset=rank * n_g, log_mean=0.05 - 0.1,
        sigma=0.2 ** 0.5, mode="log", out_data_T=buf, device=dev), 1, 3)
    del buf
    mem, batch = 1_000_000, 256
    rs = np.random.RandomState(0)
    ends = np.cumsum(rs.randint(5, 61, size=mem // 5))
    done = np.zeros(mem, dtype=bool)
    done[ends[ends < mem] - 1] = True
    rep = {}
    times = [t]
    for nstep in (1, 5, 10):
        inputs = {"gpu": str(dev), "input_dims": (5,), "num_actions": 1, "mini_batch_size": batch, "discount": 0.99,
                  "multi_steps": nstep, "r_abs_zero": None, "dynamics": "M", "buffer": mem, "n_cumsteps": mem}
        rb = ReplayBufferTorch(inputs)
        st = torch.randn((mem, 5), dtype=torch.float64, device=dev)
        rb.store_batch(st, st[:, :1], 1 + 0.01 * st[:, 0], st, torch.as_tensor(done, device=dev))
        t1 = _event_time(torch, lambda: rb.sample_exp(), 5, 50)
        tk = _event_time(torch, lambda: rb.sample_many(1024), 2, 10)
        rep[nstep] = (t1, tk)
        times += [t1, tk]
        del rb, st
    env = envs.Coin_InvA(1, n_envs=1, seed=1, device=dev)
    col = collector.Collector(env, 100_000, {"mini_batch_size": batch, "discount": 0.99, "multi_steps": 5,
                                             "r_abs_zero": None, "dynamics": "M"}, seed=2)
    act = torch.full((1, 1), 0.4, dtype=torch.float64, device=dev)
    run = col.capture(act, k=1)
    tc = _event_time(torch, run, 20, 2000)
    times.append(tc)
    if world > 1:
        tt = torch.tensor(times, dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times = [float(x) for x in tt]
    t = times[0]
    out["gbm_philox_sweep"] = {
        "value": world * n_g * h / t, "unit": UNIT, "ms_per_launch": t * 1e3,
        "workload": "lev/gbm.py final-time sweep, Philox4x32-10 + Box-Muller on device, 1.25e7 investors x 1e4 steps "
                    "per GPU, 10 leverages",
    }
    out["replay_nstep_sampling"] = {
        "unit": "samples/s", "buffer": mem, "batch": batch, "per_gpu": True,
        "per_call": {str(n): batch / times[1 + 2 * i] for i, n in enumerate((1, 5, 10))},
        "per_call_us": {str(n): times[1 + 2 * i] * 1e6 for i, n in enumerate((1, 5, 10))},
        "1024_batches_per_launch": {str(n): 1024 * batch / times[2 + 2 * i] for i, n in enumerate((1, 5, 10))},
    }
    out["collector_graph_coin_invA"] = {
        "unit": "env-steps/s", "n_envs": 1, "value": 1 / times[-1], "us_per_step_store_sample": times[-1] * 1e6,
        "note": "one CUDA-graph replay = env step + replay append + 5-step sample of 256",
    }
    return out


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from rlmd_b200 import engine, lev_exp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (rlmd_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    peaks, peak_kind = load_peaks()

    n, h = args.investors, HORIZON
    lev = np.asarray(lev_exp.param_range(*GRID), dtype=np.float32)
    table = lev_exp.dice_factor_table(lev, *RETURNS)
    g = table.shape[0]
    top_total = TOP * world
    n_total = n * world

    # synthetic outcomes of this rank's investors, resident in HBM (10 GB >> 126 MB L2:
    # every step streams the whole array from DRAM again)
    packed = args.format == "packed2"
    outcomes = engine.lev_draw("discrete", n, h, seed=420, investor_offset=rank * n, probs=PROBS, device=dev,
                               packed=packed)
    row_bytes = (h + 3) // 4 if packed else h            # algorithmic bytes per investor row
    resident = outcomes.data if packed else outcomes
    # the public pipeline object with ONE data_T buffer: every sweep waits for the previous step's statistics,
    # so a step is sweep -> statistics, one after the other (--pipeline: two buffers, the statistics of step i
    # run beside the sweep of step i+1)
    pipe = engine.FinalSweepPipeline("discrete", table, V0, top_total, device=dev, group=group, n_total=n_total,
                                     depth=2 if args.pipeline else 1)
    pipe.timing = True
    stats_holder = {}
    ev = [None] * args.steps

    def step(i=None):
        stats_holder["s"] = pipe.submit(outcomes)
        if i is not None:
            ev[i] = pipe.last_sweep

    for _ in range(max(args.warmup, 3)):
        step()
    pipe.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    cur = torch.cuda.current_stream()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step(i)
    cur.wait_stream(pipe.sweep_stream)
    cur.wait_stream(pipe.stats_stream)
    t_end.record()
    sampler.sample()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    elapsed = t_start.elapsed_time(t_end) * 1e-3
    sweep_s = sum(a.elapsed_time(b) for a, b in ev) * 1e-3 / args.steps
    if world > 1:
        t = torch.tensor([elapsed, sweep_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, sweep_s = float(t[0]), float(t[1])
    value = n_total * h * args.steps / elapsed
    # the count kernel alone (nothing beside it), after the timed region: beside the statistics kernels of the
    # previous step its live launch duration above is longer than its own
    alone_s = _event_time(torch, lambda: engine.lev_sweep("discrete", table, V0, outcomes=outcomes, mode="log",
                                                          out_data_T=pipe.data_T[0]), 2, 20)
    if world > 1:
        t = torch.tensor([alone_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        alone_s = float(t[0])
    # sweep + 4 statistic passes + 4 row-resolve kernels; multi-GPU: + 4 flag kernels (peer-memory exchange,
    # the default) or + pack / unpack around each of the 4 all-reduces (RLMD_B200_EXCHANGE=nccl)
    exchange = os.environ.get("RLMD_B200_EXCHANGE", "p2p") if world > 1 else None
    launches_per_step = 1 + 8 + (0 if world == 1 else 4 if exchange == "p2p" else 8)

    # ---- end to end: pinned host outcomes -> H2D (overlapped with the sweep) -> statistics -> host
    def measure_e2e(as_packed):
        """Pinned host outcomes -> lev_final_host (H2D inside) -> statistics on the host, CUDA-event timed."""
        n_host = n
        # packed rows keep the engine's 16-byte row padding (2512 bytes for 1e4 rolls): a chunk is then one
        # contiguous copy and every device row starts on a 16-byte boundary
        width = resident.shape[1] if (as_packed and packed) else (-(-((h + 3) // 4) // 16) * 16 if as_packed else h)
        while True:   # halve the e2e sample if the host refuses to pin that much
            try:
                host = torch.empty((n_host, width), dtype=torch.uint8, pin_memory=True)
                break
            except RuntimeError:
                n_host //= 2
                if n_host < 1000:
                    raise
        if world > 1:   # every rank must run the same sample size (the statistics are collective)
            t = torch.tensor([n_host], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            n_host = int(t[0])
            host = host[:n_host]
        if as_packed == packed:
            host.copy_(resident[:n_host, :width])
        elif as_packed:
            host.copy_(engine.pack_codes(outcomes[:n_host]).data[:, :width])
        else:
            host.copy_(outcomes.unpack()[:n_host] if n_host == n else
                       engine.PackedCodes(outcomes.data[:n_host], h).unpack())
        torch.cuda.synchronize()
        src = engine.PackedCodes(host, h) if as_packed else host
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        kw = dict(mode="log", device=dev, group=group, n_total=n_host * world)
        engine.lev_final_host("discrete", table, V0, top_total, src, **kw)  # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record()
        for _ in range(e2e_steps):
            st = engine.lev_final_host("discrete", table, V0, top_total, src, **kw)   # ends with the D2H read
        e_end.record()
        torch.cuda.synchronize()
        dt = e_start.elapsed_time(e_end) * 1e-3
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return {
            "value": n_host * world * h * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": n_host * width,
            "investors_per_gpu": n_host, "d2h_bytes_per_step": int(st.nbytes), "steps": e2e_steps,
            "host_format": "2-bit packed codes" if as_packed else "uint8 codes",
            "note": "per-rank pinned host outcomes copied H2D in 256 MiB row chunks overlapped with the sweep; "
                    "statistics read back to the host every step (PCIe-bound)",
        }

    # ---- the same K steps through a two-buffer pipeline (statistics of step i beside the sweep of step i+1)
    pipelined = None
    if not args.pipeline and not args.no_secondary:
        p2 = engine.FinalSweepPipeline("discrete", table, V0, top_total, device=dev, group=group, n_total=n_total,
                                       depth=2)
        for _ in range(3):
            p2.submit(outcomes)
        p2.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            p2.submit(outcomes)
        cur.wait_stream(p2.sweep_stream)
        cur.wait_stream(p2.stats_stream)
        b.record()
        torch.cuda.synchronize()
        dtp = a.elapsed_time(b) * 1e-3
        if world > 1:
            t = torch.tensor([dtp], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtp = float(t[0])
        pipelined = {"value": n_total * h * args.steps / dtp, "unit": UNIT, "ms_per_step": 1e3 * dtp / args.steps,
                     "note": "engine.FinalSweepPipeline with two data_T buffers, measured after the timed region"}
        del p2

    # ---- end to end: pinned host outcomes -> H2D (overlapped with the sweep) -> statistics -> host
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(packed)
        if packed:   # the same call on one-byte-per-roll host outcomes, for comparison (4x the PCIe bytes)
            other = measure_e2e(False)
            e2e["uint8_host_codes"] = {k: other[k] for k in ("value", "h2d_bytes_per_step", "investors_per_gpu")}

    secondary = None if args.no_secondary else run_secondary(torch, dist, engine, lev_exp, np, dev, rank, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic bytes: one read of this GPU's outcome array per launch (1/4 B per investor-step packed, 1 B as uint8)
    kernel = "log_discrete_packed_kernel" if packed else "log_discrete_stream_kernel"
    achieved = n * row_bytes / sweep_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and n == N_INVESTORS:
        try:
            traffic = json.load(open(tpath)).get(kernel, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": kernel + "<3>", "achieved": achieved, "peak": hbm_peak,
        "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
        "algorithmic_bytes_per_launch": n * row_bytes, "bytes_per_investor_step": row_bytes / h,
        "avg_launch_ms": sweep_s * 1e3,
        "alone": {"avg_launch_ms": alone_s * 1e3, "achieved": n * row_bytes / alone_s / 1e9,
                  "frac": n * row_bytes / alone_s / 1e9 / hbm_peak,
                  "note": "the same launch, 20 times after the timed region (with --pipeline the statistics kernels "
                          "of the previous step share the GPU with it inside the timed region)"},
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        n_s = args.cpu_sample
        dt = cpu_reference_step(n_s, h)
        cpu = {
            "value": n_s * h / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_s} investors x {h} steps, 20 leverages: torch-CPU port of dice_fixed_final_lev "
                      f"(oracle/lev_ref_port.py), {dt:.1f} s",
        }

    stats = stats_holder["s"].cpu().numpy()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("2-bit codes" if packed else "u8 codes") + ", int32 counts, f64 log-wealth",
        "data": "synthetic (on-device Philox4x32-10 die rolls, seed 420)",
        "config": {
            "workload": WORKLOAD, "investors_per_gpu": n, "horizon": h, "leverages": g, "top": top_total,
            "mode": "log-domain final sweep + exact row statistics", "sharding": f"investors x{world}",
            "pipeline": "engine.FinalSweepPipeline: the statistics of step i run beside the sweep of step i+1 "
                        "(two streams, two data_T buffers)" if args.pipeline else
                        "none: each sweep waits for the previous step's statistics",
            "statistics_exchange": {None: "none (one GPU)", "p2p": "resolve kernels sum the peers' histograms over "
                                    "NVLink peer memory", "nccl": "packed NCCL all-reduce per pass"}[exchange],
            "outcome_format": "packed 2-bit codes (2.5 GB per GPU)" if packed else "uint8 codes (10 GB per GPU)",
            "l2": f"inputs ({n * row_bytes / 1e9:.1f} GB per GPU) exceed the 126 MB L2; no explicit flush",
        },
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        "path_steps_per_s": value * g, "pipelined": pipelined, "secondary": secondary,
        "check": {"median_wealth_lev0": float(stats[0, 9]), "mean_wealth_lev0": float(stats[0, 0])},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--investors", type=int, default=N_INVESTORS, help="investors per GPU")
    ap.add_argument("--format", default="packed2", choices=["packed2", "u8"],
                    help="resident outcome format: 2-bit packed codes (default) or one uint8 per roll")
    ap.add_argument("--pipeline", action="store_true",
                    help="timed region: the statistics of step i beside the sweep of step i+1 (two streams); by default "
                         "steps run strictly one after the other and the pipelined rate is reported as `pipelined`")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the GBM-Philox / replay / collector side numbers")
    ap.add_argument("--cpu-sample", type=int, default=60_000, help="investors in the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=4_000, help="investors per step of --impl reference")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
