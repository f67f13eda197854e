"""
Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2`): two NCCL ranks,
each holding half of the investors, must return the single-GPU statistics, and
the drop-in shims must return the reference fixture's `data` from sharded rows.
On one GPU: the CUDA passes against the CPU restatement of the pass structure,
word for word on the exchanged workspace regions.
"""
import os
import socket

import numpy as np
import pytest
import torch

import golden_io
from oracle import rowstats_passes as rp
from test_oracle_lev import assert_stats_close

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_cuda_passes_fill_the_exchange_words_like_the_restatement():
    from rlmd_b200 import engine, sharding
    from rlmd_b200._lib import check, lib, ptr, stream_ptr

    rs = np.random.RandomState(5)
    rows, n, top = 3, 20_011, 13
    v = np.exp(rs.standard_normal((rows, n)) * 2).astype(np.float32)
    v[1, ::5] = v[1, 2]
    vals = torch.as_tensor(v, device="cuda")
    ws = engine.rowstats_workspace(rows, vals.device)
    stats = torch.empty((rows, 12), dtype=torch.float64, device="cuda")
    words = sharding.exchange_words(0)[4]
    ws_cpu = np.zeros((rows, words), dtype=np.int64)
    offs = {"h1": sharding.exchange_words(0)[0], "h2": sharding.exchange_words(1)[0],
            "cnt": sharding.exchange_words(2)[0], "h3": sharding.exchange_words(2)[0] + 8}
    emu = rp.RowStatsPasses(v, n, top, ws_cpu, offs)
    for phase in range(sharding.N_PHASES):
        check(lib.b200_rowstats(ptr(vals), rows, n, n, n, top, ptr(ws), ptr(stats), phase, stream_ptr()))
        emu.run(phase)
        io, ic, do, dc, _ = sharding.exchange_words(phase)
        got = ws.cpu().numpy()
        if ic:
            assert np.array_equal(got[:, io:io + ic], ws_cpu[:, io:io + ic]), f"phase {phase}: integer words"
        if dc:
            np.testing.assert_allclose(got.view(np.float64)[:, do:do + dc], ws_cpu.view(np.float64)[:, do:do + dc],
                                       rtol=1e-12)
    np.testing.assert_allclose(stats.cpu().numpy(), emu.stats, rtol=1e-12)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from rlmd_b200 import engine, lev_exp, sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        # (1) rowstats over investor shards
        rs = np.random.RandomState(21)
        v = np.exp(rs.standard_normal((5, 100_003)) * 3).astype(np.float32)
        v[2, ::9] = v[2, 1]
        off, cnt = sharding.shard_range(v.shape[1], world, rank)
        local = torch.as_tensor(v[:, off:off + cnt], device="cuda")
        st = engine.rowstats(local, 11, n_total=v.shape[1], group=dist.group.WORLD)                 # peer memory
        np.save(os.path.join(out_dir, f"stats{rank}.npy"), st.cpu().numpy())
        st_nccl = engine.rowstats(local, 11, n_total=v.shape[1], group=dist.group.WORLD, exchange="nccl")
        np.save(os.path.join(out_dir, f"stats_nccl{rank}.npy"), st_nccl.cpu().numpy())
        # back-to-back calls alternate between the two peer workspaces; rows of every key
        # mode, a different row count, a rank with far fewer investors than the other
        pw = sharding.peer_workspace(1, dist.group.WORLD, local.device)
        for it in range(6):
            rsi = np.random.RandomState(100 + it)
            vi = np.exp(rsi.standard_normal((3 + it, 20_011)) * 2).astype(np.float32)
            vi[0] = rsi.standard_normal(20_011).astype(np.float32)           # sign bit set
            vi[1, 5] = np.float32(np.nan)
            if it == 5:
                vi = vi[2:3]             # one row: a rank that owns no row still takes part in the exchange
            cut = 17 if it % 2 else 12_000
            loc = torch.as_tensor(vi[:, :cut] if rank == 0 else vi[:, cut:], device="cuda")
            sti = engine.rowstats(loc, 7, n_total=vi.shape[1], group=dist.group.WORLD)
            np.save(os.path.join(out_dir, f"loop{it}_{rank}.npy"), sti.cpu().numpy())
        assert not pw.timed_out()
        # the two-stream pipeline on shards: statistics (with their cross-GPU sums) beside the next sweep
        f = np.float32([[1.25, 0.75, 1.025], [1.5, 0.5, 1.05]])
        arrays = [engine.lev_draw("discrete", 20_000, 257, seed=50 + j, investor_offset=rank * 20_000,
                                  probs=(1 / 6, 1 / 6, 2 / 3), packed=True) for j in range(4)]
        seq = [engine.rowstats(engine.lev_sweep("discrete", f, 100.0, outcomes=a, mode="log")["data_T"], 4,
                               n_total=40_000, group=dist.group.WORLD) for a in arrays]
        pipe = engine.FinalSweepPipeline("discrete", f, 100.0, 4, group=dist.group.WORLD, n_total=40_000)
        got = [pipe.submit(a) for a in arrays]
        pipe.synchronize()
        for g_, w_ in zip(got, seq):
            assert torch.equal(g_[:, 9:12], w_[:, 9:12]) and torch.allclose(g_, w_, rtol=1e-12, atol=0)
        # (1b) the tally path on shards: count tuples -> ONE exchange of bin lists -> statistics; device codes,
        #      host int64 rows (the reference's dtype), an empty shard, and the drop-in function's printed text
        case = golden_io.lev_case("dice_top5")
        oc = golden_io.draw_outcomes(case)
        levt = np.asarray(lev_exp.param_range(*case["grid"]), dtype=np.float32)
        ft = lev_exp.dice_factor_table(levt, case["up_r"], case["down_r"], case["mid_r"])
        for k, cut in enumerate((700, 0, case["n"])):
            mine = oc[:cut] if rank == 0 else oc[cut:]
            src = engine.encode_codes(mine) if k != 1 else torch.as_tensor(mine.astype(np.int64))
            st = engine.lev_final_stats(ft, case["v0"], case["top"], src, device="cuda", group=dist.group.WORLD)
            np.save(os.path.join(out_dir, f"tally{k}_{rank}.npy"), st.cpu().numpy())
        import contextlib, io
        lev_exp.set_process_group(dist.group.WORLD)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            off, cnt = sharding.shard_range(case["n"], world, rank)
            lev_exp.dice_fixed_final_lev(torch.device("cuda", rank), torch.as_tensor(oc[off:off + cnt].astype(np.int64)),
                                         case["top"], torch.tensor(case["v0"]), case["up_r"], case["down_r"],
                                         case["mid_r"], *case["grid"])
        lev_exp.set_process_group(None)
        open(os.path.join(out_dir, f"text{rank}.txt"), "w").write(buf.getvalue())
        # (2) the dice_smart_lev shim on sharded rows of the reference's fixture
        case = golden_io.lev_case("dice_top5")
        oc = golden_io.draw_outcomes(case)
        off, cnt = sharding.shard_range(case["n"], world, rank)
        lev_exp.set_process_group(dist.group.WORLD)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            data, data_T = lev_exp.dice_smart_lev("cuda", torch.as_tensor(oc[off:off + cnt].astype(np.int64)), None,
                                                  None, case["top"], case["v0"], case["up_r"], case["down_r"],
                                                  case["mid_r"], *case["grid"])
        lev_exp.set_process_group(None)
        np.save(os.path.join(out_dir, f"data{rank}.npy"), data.cpu().numpy())
        np.save(os.path.join(out_dir, f"dataT{rank}.npy"), data_T.cpu().numpy())
        # (2a) the big-brain sweep (state-dependent leverage) on sharded rows of the reference's fixture
        bcase = golden_io.bigbrain_case("coin_grid_top4")
        boc = golden_io.draw_outcomes(bcase)
        off, cnt = sharding.shard_range(bcase["n"], world, rank)
        lev_exp.set_process_group(dist.group.WORLD)
        with contextlib.redirect_stdout(io.StringIO()):
            bdata = lev_exp.coin_big_brain_lev(
                "cuda", torch.tensor(boc[off:off + cnt].astype(np.float32)), torch.tensor(cnt, dtype=torch.int32),
                torch.tensor(bcase["h"], dtype=torch.int32), bcase["top"], torch.tensor(bcase["v0"]), bcase["up_r"],
                bcase["down_r"], torch.tensor(golden_io.bigbrain_lev_factor(bcase), dtype=torch.float64),
                *bcase["stop"], *bcase["roll"])
        lev_exp.set_process_group(None)
        np.save(os.path.join(out_dir, f"bigbrain{rank}.npy"), bdata.cpu().numpy())
        # (2b) C4's shape: a GBM sweep on investor shards (on-device Philox, global investor ids) and the
        #      growth-rate summaries of the WHOLE population from the shards
        levg = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), dtype=np.float32)
        off, cnt = sharding.shard_range(60_000, world, rank)
        gb = engine.lev_sweep("gbm", levg, 100.0, n_investors=cnt, horizon=500, seed=7, investor_offset=off,
                              log_mean=0.05 - 0.1, sigma=0.2 ** 0.5, mode="log", want_log_w=True)
        gs = engine.growth_summary(gb["log_w"], 500, 100.0, data_T=gb["data_T"], quantiles=(0.05, 0.5),
                                   n_total=60_000, group=dist.group.WORLD)
        np.save(os.path.join(out_dir, f"growth{rank}.npy"), gs.cpu().numpy())
        # (3) Philox sweeps are independent of the sharding
        lev = np.asarray(lev_exp.param_range(0.1, 1.0, 0.1), dtype=np.float32)
        off, cnt = sharding.shard_range(50_000, world, rank)
        res = engine.lev_sweep("discrete", lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05), 100.0, n_investors=cnt,
                               horizon=257, mode="chain", seed=9, investor_offset=off, probs=(1 / 6, 1 / 6, 2 / 3))
        np.save(os.path.join(out_dir, f"phx{rank}.npy"), res["data_T"].cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_equal_one(tmp_path):
    import torch.multiprocessing as mp
    from rlmd_b200 import engine, lev_exp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(21)
    v = np.exp(rs.standard_normal((5, 100_003)) * 3).astype(np.float32)
    v[2, ::9] = v[2, 1]
    want = engine.rowstats(torch.as_tensor(v, device="cuda"), 11).cpu().numpy()
    for r in range(world):
        for name in ("stats", "stats_nccl"):
            got = np.load(tmp_path / f"{name}{r}.npy")
            assert np.array_equal(got[:, 9:12], want[:, 9:12])
            np.testing.assert_allclose(got, want, rtol=1e-12)
    # the peer-memory exchange sums the ranks in one order everywhere: identical bits on both ranks
    assert np.array_equal(np.load(tmp_path / "stats0.npy"), np.load(tmp_path / "stats1.npy"))
    for it in range(6):
        rsi = np.random.RandomState(100 + it)
        vi = np.exp(rsi.standard_normal((3 + it, 20_011)) * 2).astype(np.float32)
        vi[0] = rsi.standard_normal(20_011).astype(np.float32)
        vi[1, 5] = np.float32(np.nan)
        if it == 5:
            vi = vi[2:3]
        wi = engine.rowstats(torch.as_tensor(vi, device="cuda"), 7).cpu().numpy()
        for r in range(world):
            gi = np.load(tmp_path / f"loop{it}_{r}.npy")
            assert np.array_equal(np.isnan(gi), np.isnan(wi))
            ok = ~np.isnan(wi)
            assert np.array_equal(gi[:, 9:12][ok[:, 9:12]], wi[:, 9:12][ok[:, 9:12]])
            np.testing.assert_allclose(gi[ok], wi[ok], rtol=1e-12)
    # the tally path: both ranks bit-identical, equal to one GPU, and the printed text is the reference's
    import json
    case = golden_io.lev_case("dice_top5")
    oc = golden_io.draw_outcomes(case)
    levt = np.asarray(lev_exp.param_range(*case["grid"]), dtype=np.float32)
    ft = lev_exp.dice_factor_table(levt, case["up_r"], case["down_r"], case["mid_r"])
    one = engine.lev_final_stats(ft, case["v0"], case["top"], engine.encode_codes(oc)).cpu().numpy()
    for k in range(3):
        a, b = np.load(tmp_path / f"tally{k}_0.npy"), np.load(tmp_path / f"tally{k}_1.npy")
        assert np.array_equal(a, b)
        assert np.array_equal(a[:, 9:12], one[:, 9:12])
        np.testing.assert_allclose(a, one, rtol=1e-13)
    ref_text = json.load(open(os.path.join(golden_io.GOLDEN_DIR, "final_text.json")))["dice_top5"]
    for r in range(world):
        assert open(tmp_path / f"text{r}.txt").read().rstrip("\n") == ref_text
    from test_oracle_bigbrain import assert_bigbrain_close
    bgold = golden_io.load("bigbrain_coin_grid_top4")
    b0, b1 = np.load(tmp_path / "bigbrain0.npy"), np.load(tmp_path / "bigbrain1.npy")
    assert np.array_equal(b0.view(np.uint32), b1.view(np.uint32))           # both ranks: identical bits
    assert_bigbrain_close(b0[:, :, :, bgold["cols"]], bgold["data"])        # = the reference on the whole population
    levg = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), dtype=np.float32)
    full = engine.lev_sweep("gbm", levg, 100.0, n_investors=60_000, horizon=500, seed=7, log_mean=0.05 - 0.1,
                            sigma=0.2 ** 0.5, mode="log", want_log_w=True)
    want_g = engine.growth_summary(full["log_w"], 500, 100.0, data_T=full["data_T"], quantiles=(0.05, 0.5)).cpu().numpy()
    for r in range(world):
        got_g = np.load(tmp_path / f"growth{r}.npy")
        assert np.array_equal(got_g[:, 0], want_g[:, 0])                   # valid-run counts
        assert np.array_equal(got_g[:, 4:6], want_g[:, 4:6])               # min / max: exact
        np.testing.assert_allclose(got_g, want_g, rtol=1e-11)
    case = golden_io.lev_case("dice_top5")
    gold = golden_io.load("lev_dice_top5")
    cols = gold["cols"]
    for r in range(world):
        data = np.load(tmp_path / f"data{r}.npy")
        assert np.array_equal(data[:, 9:13, cols], gold["data"][:, 9:13]), "medians / leverage row must be exact"
        assert_stats_close(data[:, :9, cols], gold["data"][:, :9])
    data_T = np.concatenate([np.load(tmp_path / f"dataT{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(data_T.view(np.uint32), gold["data_T"].view(np.uint32))
    lev = np.asarray(lev_exp.param_range(0.1, 1.0, 0.1), dtype=np.float32)
    one = engine.lev_sweep("discrete", lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05), 100.0, n_investors=50_000,
                           horizon=257, mode="chain", seed=9, probs=(1 / 6, 1 / 6, 2 / 3))["data_T"].cpu().numpy()
    two = np.concatenate([np.load(tmp_path / f"phx{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(one.view(np.uint32), two.view(np.uint32))


def test_exchange_pack_unpack_round_trip():
    """The packed fp64 staging buffer carries integer and double words exactly."""
    from rlmd_b200 import sharding
    from rlmd_b200._lib import check, lib, ptr, stream_ptr

    rs = np.random.RandomState(1)
    rows = 7
    for what, phases in (("rowstats", 5), ("growth", 7)):
        words = sharding.exchange_words(0, what)[4]
        src = torch.from_numpy(rs.randint(0, 2 ** 40, size=(rows, words)).astype(np.int64)).cuda()
        for phase in range(phases):
            io, ic, do, dc, _ = sharding.exchange_words(phase, what)
            if ic + dc == 0:
                continue
            src_f = src.view(torch.float64)
            src_f[:, do:do + dc] = torch.from_numpy(rs.standard_normal((rows, dc))).cuda()
            staging = torch.empty((rows, ic + dc), dtype=torch.float64, device="cuda")
            check(lib.b200_exchange_pack(ptr(src), words, rows, io, ic, do, dc, ptr(staging), stream_ptr()))
            assert torch.equal(staging[:, :ic], src[:, io:io + ic].double())
            dst = torch.zeros_like(src)
            check(lib.b200_exchange_unpack(ptr(dst), words, rows, io, ic, do, dc, ptr(staging * 2), stream_ptr()))
            assert torch.equal(dst[:, io:io + ic], 2 * src[:, io:io + ic])
            assert torch.equal(dst.view(torch.float64)[:, do:do + dc], 2 * src_f[:, do:do + dc])
            untouched = torch.ones(words, dtype=torch.bool)
            untouched[io:io + ic] = False
            untouched[do:do + dc] = False
            assert int(dst[:, untouched.cuda()].abs().sum()) == 0
