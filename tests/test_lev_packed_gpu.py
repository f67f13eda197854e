"""
Packed outcomes (desc.outcome_bits = 2: four 2-bit codes per byte) through the
C ABI: the LOG sweep on the packed array must return the uint8 sweep's counts,
log-wealth and wealth BIT FOR BIT, on the reference fixtures and on ragged sizes;
b200_lev_pack / the packed Philox draw must write the layout oracle/lev_oracle.py
restates.
"""
import numpy as np
import pytest
import torch

import golden_io
from oracle import lev_oracle as lo
from test_oracle_lev import oracle_inputs

pytestmark = pytest.mark.gpu

DISCRETE = [c for c in golden_io.LEV_CASES if c["kind"] != "gbm"]


@pytest.fixture(scope="module")
def eng():
    from rlmd_b200 import engine
    return engine


@pytest.mark.parametrize("h", [1, 3, 4, 5, 15, 16, 17, 63, 64, 65, 257, 1001, 4099])
@pytest.mark.parametrize("k", [2, 3, 4])
def test_pack_writes_the_restated_layout(eng, k, h):
    rs = np.random.RandomState(h * 7 + k)
    c = rs.randint(0, k, size=(37, h)).astype(np.uint8)
    p = eng.pack_codes(eng.encode_codes(c))
    assert p.horizon == h and p.data.shape[1] % 16 == 0
    assert np.array_equal(p.data.cpu().numpy(), lo.pack_codes(c, ld=p.data.shape[1]))   # pad bits and bytes are zero
    assert np.array_equal(p.unpack().cpu().numpy(), c)


def test_pack_takes_strided_and_unaligned_rows(eng):
    rs = np.random.RandomState(3)
    buf = torch.from_numpy(rs.randint(0, 3, size=(50, 333)).astype(np.uint8)).cuda()
    view = buf[:, 5:306]                       # row base off 16-byte alignment, row stride 333
    p = eng.pack_codes(view)
    assert np.array_equal(p.unpack().cpu().numpy(), view.cpu().numpy())


@pytest.mark.parametrize("probs,h", [((0.5, 0.5), 1), ((0.5, 0.5), 333), ((1 / 6, 1 / 6, 2 / 3), 64),
                                     ((1 / 6, 1 / 6, 2 / 3), 1001), ((0.1, 0.2, 0.3, 0.4), 50)])
def test_packed_philox_draw_equals_the_uint8_draw(eng, probs, h):
    n = 777
    a = eng.lev_draw("discrete", n, h, seed=11, investor_offset=5, probs=probs)
    b = eng.lev_draw("discrete", n, h, seed=11, investor_offset=5, probs=probs, packed=True)
    assert torch.equal(b.unpack(), a.contiguous())
    assert np.array_equal(b.data.cpu().numpy(), lo.pack_codes(a.cpu().numpy(), ld=b.data.shape[1]))


@pytest.mark.parametrize("case", DISCRETE, ids=lambda c: c["name"])
def test_packed_sweep_on_the_reference_fixtures(eng, case):
    oc, f, lev, _ = oracle_inputs(case)
    k = f.shape[1]
    codes = eng.encode_codes(oc)
    a = eng.lev_sweep("discrete", f, case["v0"], outcomes=codes, mode="log", want_log_w=True, want_counts=True)
    b = eng.lev_sweep("discrete", f, case["v0"], outcomes=eng.pack_codes(codes), mode="log", want_log_w=True,
                      want_counts=True)
    assert np.array_equal(b["counts"].cpu().numpy(), lo.counts_discrete(oc, k))   # pinned to the oracle
    assert torch.equal(a["counts"], b["counts"])
    assert torch.equal(a["log_w"].view(torch.int64), b["log_w"].view(torch.int64))
    assert torch.equal(a["data_T"].view(torch.int32), b["data_T"].view(torch.int32))


@pytest.mark.parametrize("k", [2, 3, 4])
@pytest.mark.parametrize("n,h", [(1, 1), (3, 7), (65, 63), (1000, 64), (513, 4097), (40, 20001)])
def test_packed_sweep_ragged_sizes(eng, k, n, h):
    rs = np.random.RandomState(n + h + k)
    c = rs.randint(0, k, size=(n, h)).astype(np.uint8)
    f = (1.0 + 0.1 * rs.standard_normal((5, k))).astype(np.float32).clip(0.5, 1.5)
    p = eng.pack_codes(eng.encode_codes(c))
    res = eng.lev_sweep("discrete", f, 100.0, outcomes=p, mode="log", want_log_w=True, want_counts=True)
    assert np.array_equal(res["counts"].cpu().numpy(), lo.counts_discrete(c, k))
    want = lo.log_wealth_discrete(c, f, 100.0)
    assert np.allclose(res["log_w"].cpu().numpy(), want, rtol=1e-13, atol=1e-11)


def test_packed_rows_off_alignment_and_garbage_pad_bits(eng):
    """Rows that start off 16-byte alignment take the byte path at both ends; pad BITS are masked, pad bytes unread."""
    rs = np.random.RandomState(9)
    n, h, k = 300, 1003, 3
    c = rs.randint(0, k, size=(n, h)).astype(np.uint8)
    nb = (h + 3) // 4
    host = np.full((n, nb + 40), 0xFF, dtype=np.uint8)             # garbage everywhere
    host[:, 3:3 + nb] = lo.pack_codes(c)
    host[:, 3 + nb - 1] |= np.uint8(0xC0)                          # h % 4 == 3: the top code of the last byte is pad
    buf = torch.from_numpy(host).cuda()
    p = eng.PackedCodes(buf[:, 3:3 + nb], h)
    f = np.float32([[1.5, 0.5, 1.05], [1.1, 0.9, 1.01]])
    res = eng.lev_sweep("discrete", f, 100.0, outcomes=p, mode="log", want_counts=True)
    assert np.array_equal(res["counts"].cpu().numpy(), lo.counts_discrete(c, k))


def test_packed_outcomes_from_host_memory(eng):
    rs = np.random.RandomState(4)
    n, h, top = 5000, 301, 3
    c = rs.randint(0, 3, size=(n, h)).astype(np.uint8)
    f = np.float32([[1.25, 0.75, 1.025], [1.5, 0.5, 1.05]])
    dev_codes = eng.encode_codes(c)
    want = eng.rowstats(eng.lev_sweep("discrete", f, 100.0, outcomes=dev_codes, mode="log")["data_T"], top)
    host = eng.PackedCodes(eng.pack_codes(dev_codes).data.cpu().pin_memory(), h)
    got = eng.lev_final_host("discrete", f, 100.0, top, host, chunk_rows=1024)
    assert np.array_equal(got, want.cpu().numpy())


def test_chain_mode_refuses_packed_outcomes(eng):
    from rlmd_b200._lib import B200Error

    p = eng.pack_codes(eng.encode_codes(np.zeros((4, 8), dtype=np.uint8)))
    with pytest.raises((B200Error, ValueError)):
        eng.lev_sweep("discrete", np.float32([[1.0, 1.1]]), 1.0, outcomes=p, mode="chain")
    with pytest.raises(ValueError):
        eng.PackedCodes(torch.zeros((4, 1), dtype=torch.uint8), 8)      # 1 byte cannot hold 8 codes


def test_pipelined_final_sweeps_equal_the_sequential_calls(eng):
    """FinalSweepPipeline overlaps the statistics of sweep i with sweep i+1: same numbers, any order of completion."""
    rs = np.random.RandomState(2)
    n, h, top = 30_000, 257, 3
    f = np.float32([[1.25, 0.75, 1.025], [1.5, 0.5, 1.05], [1.05, 0.95, 1.005]])
    arrays = [eng.pack_codes(eng.encode_codes(rs.randint(0, 3, size=(n, h)).astype(np.uint8))) for _ in range(5)]
    arrays.append(eng.encode_codes(rs.randint(0, 3, size=(n, h)).astype(np.uint8)))    # uint8 codes too
    want = [eng.rowstats(eng.lev_sweep("discrete", f, 100.0, outcomes=a, mode="log")["data_T"], top) for a in arrays]
    pipe = eng.FinalSweepPipeline("discrete", f, 100.0, top)
    got = [pipe.submit(a) for a in arrays]
    pipe.synchronize()
    for g, w in zip(got, want):
        assert torch.equal(g[:, 9:12], w[:, 9:12])                      # order statistics: exact
        assert torch.allclose(g, w, rtol=1e-12, atol=0)                 # fp64 sums: block order varies


def test_grid_of_any_size_from_one_pass_over_the_outcomes(eng):
    """
    BASELINE's 2-D grid of lev/dice_roll_sh.py (leverage x insurance fraction): 9 x 8 = 72 grid points
    (> one 64-point tile) from ONE count pass + b200_lev_from_counts per tile, against the oracle's chain
    and log-wealth for the same factor table, on uint8 and on packed outcomes; the first 64 points are
    bit-identical to the LOG sweep itself; the line b = 1 - a is the reference's dice_sh factor.
    """
    from rlmd_b200 import lev_exp

    rs = np.random.RandomState(12)
    n, h, top = 3001, 301, 2
    c = rs.choice(3, size=(n, h), p=[1 / 6, 1 / 6, 2 / 3]).astype(np.uint8)
    a, b = np.linspace(0.0, 1.2, 9, dtype=np.float32), np.linspace(0.0, 0.3, 8, dtype=np.float32)
    r, sh = (0.5, -0.5, 0.05), (-1.0, 5.0, -1.0)
    table = lo.general_factors(np.repeat(a, 8), np.tile(b, 9), r, sh)
    assert table.shape == (72, 3) and np.array_equal(table, lev_exp.grid2d_factor_table(np.repeat(a, 8), np.tile(b, 9), r, sh))
    codes = eng.encode_codes(c)
    for oc in (codes, eng.pack_codes(codes)):
        res = eng.lev_grid_sweep(table, 100.0, oc, want_log_w=True)
        assert np.array_equal(res["counts"].cpu().numpy(), lo.counts_discrete(c, 3))
        want = lo.log_wealth_discrete(c, table, 100.0)
        got = res["log_w"].cpu().numpy()
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin) and np.allclose(got[fin], want[fin], rtol=1e-13, atol=1e-11)
        chain = lo.chain_discrete(c, table, 100.0).astype(np.float64)
        wt = res["data_T"].cpu().numpy().astype(np.float64)
        ok = np.isfinite(chain) & (chain > 1e-30)
        assert (np.abs(wt[ok] - chain[ok]) / chain[ok]).max() <= 2e-5
        direct = eng.lev_sweep("discrete", table[:64], 100.0, outcomes=oc, mode="log", want_log_w=True)
        assert torch.equal(direct["data_T"].view(torch.int32), res["data_T"][:64].view(torch.int32))
        assert torch.equal(direct["log_w"].view(torch.int64), res["log_w"][:64].view(torch.int64))
        again = eng.lev_grid_sweep(table[5:70], 100.0, oc, counts=res["counts"])      # reuse the counts
        assert torch.equal(again["data_T"].view(torch.int32), res["data_T"][5:70].view(torch.int32))
    stats, data_T = lev_exp.dice_sh_grid2d("cuda", torch.as_tensor(c.astype(np.int64)), top, 100.0, *r, *sh, a, b)
    assert tuple(stats.shape) == (9, 8, 12) and tuple(data_T.shape) == (9, 8, n)
    want_stats = np.stack([lo.summary_stats(res["data_T"][g].cpu().numpy(), top) for g in range(72)]).reshape(9, 8, 12)
    assert np.array_equal(stats.cpu().numpy()[..., 9:12], want_stats[..., 9:12])
    assert np.allclose(stats.cpu().numpy()[..., :9], want_stats[..., :9], rtol=1e-10)
    # the same grid from the tally of count tuples (no data_T): identical order statistics, fp64 moments
    for oc in (torch.as_tensor(c.astype(np.int64)), eng.pack_codes(codes)):
        st2, none = lev_exp.dice_sh_grid2d("cuda", oc, top, 100.0, *r, *sh, a, b, want_data_T=False)
        assert none is None and tuple(st2.shape) == (9, 8, 12)
        assert torch.equal(st2[..., 9:12], stats[..., 9:12])
        assert torch.allclose(st2, stats, rtol=1e-12, atol=0)
    # the reference's own 1-D dice_sh grid is the line b = 1 - a of this family
    lev = np.float32([0.73, 0.85, 1.0])
    line = lev_exp.grid2d_factor_table(lev, (np.float32(1) - lev).astype(np.float32), r, sh)
    assert np.array_equal(line, lev_exp.dice_sh_factor_table(lev, *r, *sh))


def test_one_bit_coin_format(eng):
    """The coin at one bit per flip: pack / draw / LOG sweep / tally agree with the uint8 route, ragged sizes."""
    from rlmd_b200 import lev_exp

    rs = np.random.RandomState(3)
    lev = lo.lev_grid(0.1, 1.0, 0.1, 0.5, -0.4)
    f = lo.coin_factors(lev, 0.5, -0.4)
    for n, h in ((2001, 3000), (37, 41), (5, 8), (130, 1)):
        c = (rs.random_sample((n, h)) < 0.5).astype(np.uint8)
        codes = eng.encode_codes(c)
        p1 = eng.pack_codes(codes, bits=1)
        assert p1.bits == 1 and p1.data.shape[1] % 16 == 0 and p1.data.shape[1] * 8 >= h
        assert np.array_equal(p1.unpack().cpu().numpy(), c)
        # bit t & 7 of byte t >> 3
        want = np.packbits(c, axis=1, bitorder="little")
        assert np.array_equal(p1.data.cpu().numpy()[:, : want.shape[1]], want)
        a = eng.lev_sweep("discrete", f, 100.0, outcomes=codes, mode="log", want_log_w=True, want_counts=True)
        b = eng.lev_sweep("discrete", f, 100.0, outcomes=p1, mode="log", want_log_w=True, want_counts=True)
        assert torch.equal(a["counts"], b["counts"]) and torch.equal(a["log_w"], b["log_w"])
        assert torch.equal(a["data_T"].view(torch.int32), b["data_T"].view(torch.int32))
        assert np.array_equal(b["counts"].cpu().numpy(), lo.counts_discrete(c, 2))
        if n > 4:
            s_a = eng.lev_final_stats(f, 100.0, 2, codes).cpu().numpy()
            s_b = eng.lev_final_stats(f, 100.0, 2, p1).cpu().numpy()
            assert np.array_equal(s_a[:, 9:12], s_b[:, 9:12]) and np.allclose(s_a, s_b, rtol=1e-13, equal_nan=True)
    # the Philox draw in the one-bit format is the uint8 draw, packed
    d8 = eng.lev_draw("discrete", 1000, 333, seed=9, probs=(0.4, 0.6))
    d1 = eng.lev_draw("discrete", 1000, 333, seed=9, probs=(0.4, 0.6), packed=True, bits=1)
    assert torch.equal(d1.unpack(), d8)
    with pytest.raises(Exception):
        eng.lev_draw("discrete", 10, 10, seed=1, probs=(1 / 6, 1 / 6, 2 / 3), packed=True, bits=1)
    with pytest.raises(Exception):      # three outcomes do not fit one bit
        eng.lev_sweep("discrete", lo.dice_factors(lev, 0.5, -0.5, 0.05), 100.0, outcomes=d1, mode="log")
