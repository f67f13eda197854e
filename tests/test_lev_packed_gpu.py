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
