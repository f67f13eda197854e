"""
The per-step series path (`*_smart_lev`'s data[L,13,H-1]) on investor shards: one process per GPU
(torchrun), 1e6 investors per GPU, the dice gamble on an on-device Philox stream; every chunk of
steps is dumped, reduced by b200_rowstats_p2p (the cross-GPU sums over peer memory) and scattered.
Prints one JSON line from rank 0: path-steps/s of the whole job, ms per call, max over ranks.

  python tools/bench_series_n.py                                      # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_series_n.py
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--investors", type=int, default=1_000_000)
    ap.add_argument("--horizon", type=int, default=512)
    ap.add_argument("--grid", type=int, default=10)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    a = ap.parse_args()
    from rlmd_b200 import engine, lev_exp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    lev = np.linspace(0.05, 0.5, a.grid).astype(np.float32)
    f = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
    n = a.investors

    def call():
        return engine.lev_series("discrete", f, lev, 100.0, max(2, n * world // 1000), n_investors=n, horizon=a.horizon,
                                 seed=3, investor_offset=rank * n, probs=(1 / 6, 1 / 6, 2 / 3), n_total=n * world,
                                 group=group, device=torch.device("cuda", local))

    os.environ["RLMD_B200_EXCHANGE"] = a.exchange
    data, _ = call()
    torch.cuda.synchronize()
    times = []
    for _ in range(a.iters):
        if group is not None:
            dist.barrier(group=group)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        data, _ = call()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        times.append(float(t.item()))
    ms = float(np.median(times))
    if rank == 0:
        print(json.dumps({"what": "series", "n_gpus": world, "investors_per_gpu": n, "horizon": a.horizon,
                          "grid": a.grid, "exchange": a.exchange, "ms_per_call": ms, "ms_all": times,
                          "path_steps_per_s": world * n * a.horizon * a.grid / (ms * 1e-3),
                          "checksum": float(data[:, 9, -1].double().sum().item())}), flush=True)
    if group is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
