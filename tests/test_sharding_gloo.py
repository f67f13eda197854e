"""
The multi-GPU path on CPU: two gloo ranks run rlmd_b200.sharding.exchange_phases
with oracle/rowstats_passes.py standing in for the CUDA passes (same workspace
words, as named by b200_rowstats_exchange) on investor SHARDS, and must arrive at
the reference's sort-based statistics of the WHOLE vector on both ranks.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lev_oracle as lo
from oracle import rowstats_passes as rp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _vectors(rows=4, n=1501):
    rs = np.random.RandomState(11)
    v = np.exp(rs.standard_normal((rows, n)) * 3).astype(np.float32)
    v[1, ::7] = v[1, 3]                 # heavy ties, also across the top-K threshold
    v[2, :5] = np.float32(np.inf)       # overflowed investors: torch.std_mean -> nan
    v[3] = np.float32(0.0)              # fully underflowed row
    v[3, :9] = np.float32(1e-42)        # denormals
    return v


def _worker(rank, world, port, top, out_dir):
    from rlmd_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        v = _vectors()
        rows, n_total = v.shape
        off, cnt = sharding.shard_range(n_total, world, rank)
        assert sharding.global_count(cnt, dist.group.WORLD, "cpu") == n_total
        io1, ic1, _, _, words = sharding.exchange_words(0)
        io2 = sharding.exchange_words(1)[0]
        ioc = sharding.exchange_words(2)[0]      # phase 2 exchanges the 8 count words, then hist3
        io3 = ioc + 8
        ws = torch.zeros((rows, words), dtype=torch.int64)
        passes = rp.RowStatsPasses(v[:, off:off + cnt], n_total, top, ws.numpy(),
                                   {"h1": io1, "h2": io2, "h3": io3, "cnt": ioc})
        sharding.exchange_phases(passes.run, ws, dist.group.WORLD)
        np.save(os.path.join(out_dir, f"stats{rank}.npy"), passes.stats)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,top", [(2, 1), (2, 7), (3, 150)])
def test_sharded_statistics_equal_the_global_ones(tmp_path, world, top):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, top, str(tmp_path)), nprocs=world, join=True)
    v = _vectors()
    want = np.stack([lo.summary_stats(v[r], top) for r in range(v.shape[0])])
    for rank in range(world):
        got = np.load(tmp_path / f"stats{rank}.npy")
        assert np.array_equal(np.isnan(got), np.isnan(want)), (rank, got, want)
        ok = ~np.isnan(want)
        assert np.array_equal(got[:, 9:12][ok[:, 9:12]], want[:, 9:12][ok[:, 9:12]]), "order statistics must be exact"
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-10, atol=0)


def test_shard_range_partitions_exactly():
    from rlmd_b200 import sharding

    for n, w in [(10, 3), (1_000_000, 8), (5, 8), (0, 2), (100_000_001, 8)]:
        spans = [sharding.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(w - 1))
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
