"""
Parity at BASELINE.json's FULL sizes, where the CPU oracle cannot run, through
size-independent properties of the path (the oracle pins the same kernels on
small inputs in the other files):

* C2  dice, 1e6 investors x 1e4 steps: outcome counts are a checksum of the row
  (sum = H) and follow the die's probabilities; log-wealth is the linear form
  log V0 + sum_k n_k log m_gk of those counts (recomputed in torch fp64); the exact
  fp32 chain agrees with exp(log-wealth); the three chain variants are
  bit-identical; the 2-bit packed array gives the uint8 sweep bit for bit;
  investor slices reproduce the full run; Philox shards are invisible; the
  order statistics equal torch.sort's.
* C4  GBM, one GPU's shard (1.25e7 x 1e4, on-device Philox): W(l) W(-l) = V0^2,
  growth rates follow l (mu - sigma^2/2), valid-run counts equal the finite
  positive entries.
* C5  replay, full 1e6 buffer, n = 10: every sampled target recomputed by torch
  from the memories.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, H, TOP, V0 = 1_000_000, 10_000, 100, 100.0
PROBS = (1 / 6, 1 / 6, 2 / 3)


@pytest.fixture(scope="module")
def c2():
    from rlmd_b200 import engine, lev_exp
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a B200-class GPU (10 GB of outcomes)")
    lev = np.asarray(lev_exp.param_range(0.10, 1.00, 0.10), dtype=np.float32)
    table = lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05)
    codes = engine.lev_draw("discrete", N, H, seed=420, probs=PROBS)
    log = engine.lev_sweep("discrete", table, V0, outcomes=codes, mode="log", want_log_w=True, want_counts=True)
    yield dict(engine=engine, lev=lev, table=table, codes=codes, log=log)
    del codes, log
    torch.cuda.empty_cache()


def test_counts_are_a_checksum_of_every_row(c2):
    counts = c2["log"]["counts"]
    assert counts.shape == (N, 3)
    assert bool((counts.sum(dim=1) == H).all())
    frac = counts.double().sum(dim=0).cpu().numpy() / (N * H)
    sigma = np.sqrt(np.asarray(PROBS) * (1 - np.asarray(PROBS)) / (N * H))
    assert np.all(np.abs(frac - PROBS) < 6 * sigma)
    # a direct recount of a few thousand rows with torch
    rows = torch.arange(0, N, 251, device="cuda")
    sub = c2["codes"][rows].long()
    want = torch.stack([(sub == k).sum(dim=1) for k in range(3)], dim=1).int()
    assert torch.equal(counts[rows], want)


def test_log_wealth_is_the_linear_form_of_the_counts(c2):
    counts, log_w = c2["log"]["counts"].double(), c2["log"]["log_w"]
    lm = torch.as_tensor(np.log(c2["table"].astype(np.float64)), device="cuda")          # [G,3]
    want = float(np.log(np.float64(np.float32(V0)))) + lm @ counts.T                       # [G,N]
    fin = torch.isfinite(want)
    assert torch.equal(torch.isfinite(log_w), fin)
    assert float((log_w[fin] - want[fin]).abs().max()) <= 1e-9
    growth = (log_w[fin] - np.log(V0)) / H
    assert float(((growth - (want[fin] - np.log(V0)) / H).abs() / growth.abs().clamp_min(1e-3)).max()) <= 1e-12


def test_chain_variants_are_bit_identical_and_agree_with_the_log_form(c2):
    eng = c2["engine"]
    outs = [eng.lev_sweep("discrete", c2["table"], V0, outcomes=c2["codes"], mode="chain", variant=v)["data_T"]
            for v in (1, 2, 3)]
    assert torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32))
    assert torch.equal(outs[0].view(torch.int32), outs[2].view(torch.int32))
    chain = outs[2].double()
    ref = torch.exp(c2["log"]["log_w"])
    ok = torch.isfinite(ref) & (ref > 1e-30) & (ref < 1e38)
    # the sweep that also writes data_T agrees with exp(log_w) to fp32 rounding
    dT = c2["log"]["data_T"].double()
    assert float(((dT[ok] - ref[ok]).abs() / ref[ok]).max()) <= 1e-6
    # the fp32 chain saturates for good once it leaves the fp32 range on the way (as the
    # reference's does): compare the paths that stayed finite; one fp32 rounding per step
    # accumulates like a random walk, a path that dipped into denormals loses more
    live = ok & torch.isfinite(chain) & (chain > 0)
    assert float(live.double().mean()) > 0.5          # the top leverages mostly leave the fp32 range at H = 1e4
    rel = (chain[live] - ref[live]).abs() / ref[live]
    tol = 1.5e-5 * (H / 300) ** 0.5
    assert float(rel.median()) <= tol / 10
    assert float((rel <= tol).double().mean()) >= 0.9999
    # the lowest leverage (10 %) never leaves the range: every path within the tolerance
    assert bool(live[0].all()) and float(((chain[0] - ref[0]).abs() / ref[0]).max()) <= tol


def test_packed_array_gives_the_same_sweep_at_full_size(c2):
    """
    The bench's format: the 1e6 x 1e4 die rolls at 2 bits each (2.5 GB).  Packing the
    uint8 draws and drawing straight into the packed layout give the same bytes, and
    the sweep over them returns the uint8 sweep's counts, log-wealth and wealth bit for bit.
    """
    eng = c2["engine"]
    packed = eng.pack_codes(c2["codes"])
    drawn = eng.lev_draw("discrete", N, H, seed=420, probs=PROBS, packed=True)
    assert packed.data.shape == (N, 2512) and torch.equal(packed.data, drawn.data)
    got = eng.lev_sweep("discrete", c2["table"], V0, outcomes=packed, mode="log", want_log_w=True, want_counts=True)
    ref = c2["log"]
    assert torch.equal(got["counts"], ref["counts"])
    assert torch.equal(got["log_w"].view(torch.int64), ref["log_w"].view(torch.int64))
    assert torch.equal(got["data_T"].view(torch.int32), ref["data_T"].view(torch.int32))
    # and the statistics of the 20-leverage script grid through the two-stream pipeline equal the direct call
    want = eng.rowstats(ref["data_T"], TOP)
    pipe = eng.FinalSweepPipeline("discrete", c2["table"], V0, TOP)
    st = [pipe.submit(packed) for _ in range(3)]
    pipe.synchronize()
    for s_ in st:
        assert torch.equal(s_[:, 9:12], want[:, 9:12])                  # order statistics: exact
        # fp64 sums of 1e6 heavy-tailed terms, block order not fixed (atomics): last digits differ
        assert torch.allclose(s_, want, rtol=1e-9, atol=0, equal_nan=True)


def test_investor_slices_reproduce_the_full_run(c2):
    eng = c2["engine"]
    full = eng.lev_sweep("discrete", c2["table"], V0, outcomes=c2["codes"], mode="chain")["data_T"]
    for lo, hi in ((0, 1), (128, 257), (499_999, 500_130), (N - 77, N)):
        part = eng.lev_sweep("discrete", c2["table"], V0, outcomes=c2["codes"][lo:hi], mode="chain")["data_T"]
        assert torch.equal(part, full[:, lo:hi])
        part = eng.lev_sweep("discrete", c2["table"], V0, outcomes=c2["codes"][lo:hi], mode="log",
                             want_log_w=True)["log_w"]
        assert torch.equal(part, c2["log"]["log_w"][:, lo:hi])


def test_philox_mode_equals_the_streamed_draws_and_shards_are_invisible(c2):
    eng = c2["engine"]
    streamed = eng.lev_sweep("discrete", c2["table"], V0, outcomes=c2["codes"], mode="chain")["data_T"]
    philox = eng.lev_sweep("discrete", c2["table"], V0, n_investors=N, horizon=H, seed=420, probs=PROBS,
                           mode="chain")["data_T"]
    assert torch.equal(streamed, philox)
    off = 625_000
    shard = eng.lev_sweep("discrete", c2["table"], V0, n_investors=N - off, horizon=H, seed=420, probs=PROBS,
                          investor_offset=off, mode="chain")["data_T"]
    assert torch.equal(shard, streamed[:, off:])


def test_order_statistics_equal_torch_sort(c2):
    eng = c2["engine"]
    data_T = c2["log"]["data_T"]
    stats = eng.rowstats(data_T, TOP).cpu().numpy()
    for g in (0, 4, 9):
        s = torch.sort(data_T[g], descending=True)[0]
        top, adj = s[:TOP], s[TOP:]
        low_med = lambda v: float(torch.sort(v)[0][(v.numel() - 1) // 2])
        assert stats[g, 9] == low_med(s) and stats[g, 10] == low_med(top) and stats[g, 11] == low_med(adj)
        for j, grp in enumerate((s, top, adj)):
            d = grp.double()
            if not bool(torch.isfinite(d).all()):     # torch.std_mean semantics: one inf -> nan mean / MAD / std
                assert np.isnan(stats[g, [j, 3 + j, 6 + j]]).all()
                continue
            mean = d.mean()
            assert stats[g, j] == pytest.approx(float(mean), rel=1e-11)
            assert stats[g, 3 + j] == pytest.approx(float((d - mean).abs().mean()), rel=1e-10)
            assert stats[g, 6 + j] == pytest.approx(float(((d - mean) ** 2).mean().sqrt()), rel=1e-10)
    # idempotence: the statistics of the sorted row are the statistics of the row
    srt = torch.sort(data_T[:3], dim=1)[0].contiguous()
    assert np.array_equal(eng.rowstats(srt, TOP).cpu().numpy()[:, 9:12], stats[:3, 9:12])


def test_gbm_shard_properties():
    from rlmd_b200 import engine, lev_exp
    n, h = 12_500_000, 10_000
    mu, sg = 0.05 - 0.2 / 2, 0.2 ** 0.5
    lev = np.asarray(lev_exp.param_range(-1.0, 1.0, 0.2), dtype=np.float32)      # 10 values, symmetric, no 0
    assert len(lev) == 10 and np.allclose(lev, -lev[::-1])
    res = engine.lev_sweep("gbm", lev, V0, n_investors=n, horizon=h, seed=7, investor_offset=3 * n, log_mean=mu,
                           sigma=sg, mode="log", want_log_w=True)
    lw, dT = res["log_w"], res["data_T"]
    logv0 = np.log(V0)
    # log W(l) - log V0 = l * S: one running sum S per investor serves the whole grid
    s_est = (lw - logv0) / torch.as_tensor(lev.astype(np.float64), device="cuda")[:, None]
    assert float((s_est - s_est[-1]).abs().max()) <= 1e-8
    # S / H is the sample mean of N(mu, sigma^2 / H) draws
    s_over_h = (lw[-1] - logv0) / float(lev[-1]) / h
    assert abs(float(s_over_h.mean()) - mu) < 6 * sg / np.sqrt(n * h)
    assert abs(float(s_over_h.std()) - sg / np.sqrt(h)) < 0.01 * sg / np.sqrt(h)
    summ = engine.growth_summary(lw, h, V0, data_T=dT, quantiles=(0.05, 0.5)).cpu().numpy()
    valid = (torch.isfinite(dT) & (dT > 0)).sum(dim=1).cpu().numpy()
    assert np.array_equal(summ[:, 0], valid)
    g = (lw - logv0) / h
    assert np.allclose(summ[:, 1], g.mean(dim=1).cpu().numpy(), rtol=1e-9, atol=1e-15)
    med = torch.sort(g[2])[0]
    # type-8 median of an even count = mean of the two middle order statistics
    assert summ[2, 7] == pytest.approx(float((med[n // 2 - 1] + med[n // 2]) / 2), rel=1e-12)
    del res, lw, dT
    torch.cuda.empty_cache()


def test_replay_full_buffer_nstep_targets_recomputed_by_torch():
    from rlmd_b200.replay_torch import ReplayBufferTorch
    mem, nstep, batch, gamma = 1_000_000, 10, 512, 0.99
    rs = np.random.RandomState(0)
    ends = np.cumsum(rs.randint(5, 61, size=mem // 5))
    done = np.zeros(mem, dtype=bool)
    done[ends[ends < mem] - 1] = True
    inputs = {"gpu": "cuda:0", "input_dims": (5,), "num_actions": 1, "mini_batch_size": batch, "discount": gamma,
              "multi_steps": nstep, "r_abs_zero": None, "dynamics": "M", "buffer": mem, "n_cumsteps": mem}
    buf = ReplayBufferTorch(inputs)
    st = torch.randn((mem, 5), dtype=torch.float64, device="cuda")
    buf.store_batch(st, st[:, :1], 1 + 0.01 * st[:, 0], st + 1, torch.as_tensor(done, device="cuda"))
    idx, s, a, r, s2, d, eff = buf.sample_many(64)
    idx, eff = idx.reshape(-1), eff.reshape(-1)
    assert int(idx.min()) >= 0 and int(idx.max()) < mem
    assert all(len(torch.unique(row)) == batch for row in idx.view(64, batch))
    assert int(eff.min()) >= 1 and int(eff.max()) <= nstep
    # beyond the first episode: history = own episode up to the slot (+1 future step unless terminal)
    term = torch.as_tensor(done, device="cuda")
    e0 = int(np.flatnonzero(done)[0])
    elast = int(np.flatnonzero(done)[-1])
    inner = (idx > e0) & (idx <= elast)
    start = buf.episode_start[idx].long()
    length = idx - start + 1 + (~term[idx]).long()
    assert torch.equal(eff[inner], torch.minimum(length, torch.tensor(nstep, device="cuda"))[inner])
    first = start + length - eff
    acc = torch.ones_like(r.reshape(-1))
    for t in range(nstep - 1):
        use = inner & (t < eff - 1)
        term_t = torch.as_tensor(np.float32(gamma ** t), device="cuda") * buf.reward_memory[(first + t).clamp(0, mem - 1)]
        acc = torch.where(use, acc * term_t, acc)
    assert torch.equal(r.reshape(-1)[inner], acc[inner])
    assert torch.equal(s.reshape(-1, 5)[inner], buf.next_state_memory[first.clamp(0, mem - 1)][inner])
    assert torch.equal(s2.reshape(-1, 5), buf.next_state_memory[idx])
    assert torch.equal(d.reshape(-1), term[idx])


def test_tally_path_at_full_size(c2):
    """C2 through the route the drop-in functions take: count kernel with the tally sink -> distinct count
    tuples -> cluster select.  Size-independent properties: the bins are exactly the distinct rows of the
    count matrix and their weights add up to N; the order statistics are those of the sweep's data_T
    (torch.sort), bit for bit; the moments agree with the row-statistics path to fp64 rounding; the uint8,
    the packed and the int64 (reference dtype, ingest kernel) forms of the outcomes give the same answer."""
    from rlmd_b200 import tally
    eng = c2["engine"]
    table, codes = c2["table"], c2["codes"]
    t = tally.FinalTally(N)
    t.add(codes, 3)
    t.finalize()
    st = t.stats(table, V0, H, n_total=N, top=TOP)
    info = t.check()
    counts = c2["log"]["counts"]
    key = counts[:, 1].long() * (H + 1) + counts[:, 2].long()
    assert info["bins"] == int(torch.unique(key).numel())
    rows = eng.rowstats(c2["log"]["data_T"], TOP)
    assert torch.equal(st[:, 9:12], rows[:, 9:12])
    fin = torch.isfinite(rows)
    assert torch.equal(torch.isnan(st), torch.isnan(rows))
    assert float(((st[fin] - rows[fin]).abs() / rows[fin].abs().clamp_min(1e-300)).max()) <= 1e-12
    srt = torch.sort(c2["log"]["data_T"][3])[0]
    assert float(st[3, 9]) == float(srt[(N - 1) // 2]) and float(st[3, 10]) == float(srt[N - TOP + (TOP - 1) // 2])
    packed = eng.pack_codes(codes)
    st_p = eng.lev_final_stats(table, V0, TOP, packed)
    assert torch.equal(st_p[:, 9:12], st[:, 9:12]) and torch.allclose(st_p, st, rtol=1e-12, atol=0, equal_nan=True)
    # the reference's dtype for a slice of the investors (8 B per roll: 2e5 rows = 16 GB)
    n_s = 200_000
    wide = codes[:n_s].to(torch.int64)
    a = eng.lev_final_stats(table, V0, 20, wide)
    del wide
    b = eng.lev_final_stats(table, V0, 20, codes[:n_s])
    assert torch.equal(a[:, 9:12], b[:, 9:12]) and torch.allclose(a, b, rtol=1e-12, atol=0, equal_nan=True)
