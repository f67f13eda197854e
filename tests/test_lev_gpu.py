"""
GPU parity tests of the leverage sweep (K1) and the row statistics, through the
C ABI, against (a) the fixtures written by the unmodified reference and (b) the
CPU oracle on the same seeded inputs.
"""
import numpy as np
import pytest
import torch

import golden_io
from oracle import lev_oracle as lo
from oracle import philox_oracle as po
from test_oracle_lev import assert_stats_close, oracle_inputs

pytestmark = pytest.mark.gpu

DISCRETE = [c for c in golden_io.LEV_CASES if c["kind"] != "gbm"]
GBM = [c for c in golden_io.LEV_CASES if c["kind"] == "gbm"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def eng():
    from rlmd_b200 import engine
    return engine


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("case", DISCRETE, ids=lambda c: c["name"])
def test_chain_data_T_bit_exact_vs_reference(eng, case, variant):
    oc, f, lev, _ = oracle_inputs(case)
    gold = golden_io.load("lev_" + case["name"], oc)
    codes = eng.encode_codes(oc)
    res = eng.lev_sweep("discrete", f, case["v0"], outcomes=codes, mode="chain", variant=variant)
    got = res["data_T"].cpu().numpy()
    assert np.array_equal(bits(got), bits(gold["data_T"]))


@pytest.mark.parametrize("case", DISCRETE, ids=lambda c: c["name"])
def test_chain_unaligned_rows_take_the_plain_load_path(eng, case):
    oc, f, lev, _ = oracle_inputs(case)
    gold = golden_io.load("lev_" + case["name"], oc)
    n, h = oc.shape
    buf = torch.zeros((n, h + 3), dtype=torch.uint8, device="cuda")  # odd stride: no TMA
    buf[:, 1:h + 1] = torch.from_numpy(oc).cuda()
    res = eng.lev_sweep("discrete", f, case["v0"], outcomes=buf[:, 1:h + 1], mode="chain")
    assert np.array_equal(bits(res["data_T"].cpu().numpy()), bits(gold["data_T"]))


@pytest.mark.parametrize("case", DISCRETE, ids=lambda c: c["name"])
def test_final_statistics_vs_reference(eng, case):
    oc, f, lev, _ = oracle_inputs(case)
    gold = golden_io.load("lev_" + case["name"], oc)
    res = eng.lev_sweep("discrete", f, case["v0"], outcomes=eng.encode_codes(oc), mode="chain")
    stats = eng.rowstats(res["data_T"], case["top"]).cpu().numpy()     # [G,12]
    want = gold["data"][:, :12, -1]                                    # last kept column = step H-1
    # order statistics of bit-exact values are bit-exact
    assert np.array_equal(bits(stats[:, 9:12].astype(np.float32)), bits(want[:, 9:12]))
    assert_stats_close(stats[:, :9, None], want[:, :9, None])


@pytest.mark.parametrize("case", DISCRETE, ids=lambda c: c["name"])
def test_log_mode_counts_and_growth(eng, case):
    oc, f, lev, _ = oracle_inputs(case)
    k = f.shape[1]
    res = eng.lev_sweep("discrete", f, case["v0"], outcomes=eng.encode_codes(oc), mode="log",
                        want_log_w=True, want_counts=True)
    assert np.array_equal(res["counts"].cpu().numpy(), lo.counts_discrete(oc, k))   # bit-exact indexing
    want = lo.log_wealth_discrete(oc, f, case["v0"])
    got = res["log_w"].cpu().numpy()
    h = case["h"]
    v0 = np.log(np.float64(np.float32(case["v0"])))
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    # growth rates g = (log W - log V0) / H to 1e-12 relative (fp64 path)
    gw, gg = (want[fin] - v0) / h, (got[fin] - v0) / h
    assert np.all(np.abs(gg - gw) <= 1e-12 * np.maximum(np.abs(gw), 1e-3))
    # and the fp32 wealth agrees with the reference chain to its own rounding noise
    gold = golden_io.load("lev_" + case["name"], oc)["data_T"].astype(np.float64)
    wt = res["data_T"].cpu().numpy().astype(np.float64)
    ok = np.isfinite(gold) & (gold > 1e-30)
    rel = np.abs(wt[ok] - gold[ok]) / gold[ok]
    assert rel.max() <= 1e-5 * max(1.0, (h / 300) ** 0.5) * 1.5


@pytest.mark.parametrize("case", GBM, ids=lambda c: c["name"])
def test_gbm_vs_reference(eng, case):
    x, lev, _, _ = oracle_inputs(case)
    gold = golden_io.load("lev_" + case["name"], x)
    res = eng.lev_sweep("gbm", lev, case["v0"], outcomes=eng.encode_returns(x), mode="log", want_log_w=True)
    got = res["data_T"].cpu().numpy()
    ref = gold["data_T"]
    # valid-run pattern (finite and > 0 in the reference dtype), allowing the
    # handful of paths that sit within rounding of the fp32 range limit
    valid_ref = np.isfinite(ref) & (ref > 0)
    valid_got = np.isfinite(got) & (got > 0)
    assert (valid_ref != valid_got).mean() <= 2e-3
    both = valid_ref & valid_got & (ref > 1e-30)
    rel = np.abs(got[both].astype(np.float64) - ref[both]) / ref[both]
    assert rel.max() <= 1e-5 * max(1.0, case["h"] / 500)
    # fp64 log-wealth against the oracle's fp64 sum of fl32 products: the engine
    # uses l * sum(x); the difference is the fp32 rounding of each product
    want = lo.log_wealth_gbm(x, lev, case["v0"])
    assert np.abs(res["log_w"].cpu().numpy() - want).max() <= 2e-5 * max(1.0, case["h"] / 500) ** 0.5


def test_gbm_unaligned_rows(eng):
    case = golden_io.lev_case("gbm_snp_top4")
    x, lev, _, _ = oracle_inputs(case)
    n, h = x.shape
    a = eng.lev_sweep("gbm", lev, 100.0, outcomes=eng.encode_returns(x), mode="log", want_log_w=True)
    buf = torch.zeros((n, h + 3), dtype=torch.float32, device="cuda")
    buf[:, 1:h + 1] = torch.from_numpy(x).cuda()
    b = eng.lev_sweep("gbm", lev, 100.0, outcomes=buf[:, 1:h + 1], mode="log", want_log_w=True)
    assert torch.equal(a["log_w"], b["log_w"]) and torch.equal(a["data_T"], b["data_T"])


@pytest.mark.parametrize("probs,h", [((0.5, 0.5), 301), ((1 / 6, 1 / 6, 2 / 3), 64), ((0.1, 0.2, 0.3, 0.4), 7)])
def test_philox_discrete_outcomes_bit_exact(eng, probs, h):
    n, seed, off = 1000, 0x1234567890ABCDEF, 3_000_000_000
    drawn = eng.lev_draw("discrete", n, h, seed=seed, investor_offset=off, probs=probs)
    want = po.discrete_codes(seed, np.arange(n, dtype=np.uint64) + np.uint64(off), h, probs)
    assert np.array_equal(drawn.cpu().numpy(), want)
    k = len(probs)
    f = np.linspace(0.8, 1.3, 6 * k).reshape(6, k).astype(np.float32)
    a = eng.lev_sweep("discrete", f, 100.0, n_investors=n, horizon=h, seed=seed, investor_offset=off,
                      probs=probs, mode="chain")
    b = eng.lev_sweep("discrete", f, 100.0, outcomes=drawn, mode="chain")
    assert torch.equal(a["data_T"], b["data_T"])
    assert np.array_equal(bits(a["data_T"].cpu().numpy()), bits(lo.chain_discrete(want, f, 100.0)))


def test_philox_gbm_matches_streamed_draws(eng):
    n, h, seed = 3000, 203, 99
    mu, sg = 0.05 - 0.2 / 2, 0.2 ** 0.5
    lev = lo.lev_grid(-1.0, 1.0, 0.2)
    x = eng.lev_draw("gbm", n, h, seed=seed, log_mean=mu, sigma=sg)
    a = eng.lev_sweep("gbm", lev, 100.0, n_investors=n, horizon=h, seed=seed, log_mean=mu, sigma=sg,
                      mode="log", want_log_w=True)
    b = eng.lev_sweep("gbm", lev, 100.0, outcomes=x, mode="log", want_log_w=True)
    assert torch.equal(a["log_w"], b["log_w"]) and torch.equal(a["data_T"], b["data_T"])
    want = po.gbm_returns(seed, np.arange(n), h, mu, sg)
    assert np.abs(x.cpu().numpy() - want).max() < 2e-5
    z = (x.cpu().numpy().astype(np.float64) - mu) / sg
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3


def test_philox_sharding_is_invisible(eng):
    """Two half-shards with investor_offset reproduce the single run exactly."""
    n, h = 1024, 50
    f = lo.dice_factors(lo.lev_grid(0.1, 1.0, 0.1), 0.5, -0.5, 0.05)
    kw = dict(horizon=h, seed=5, probs=(1 / 6, 1 / 6, 2 / 3), mode="chain")
    full = eng.lev_sweep("discrete", f, 100.0, n_investors=n, **kw)["data_T"]
    lo_half = eng.lev_sweep("discrete", f, 100.0, n_investors=n // 2, **kw)["data_T"]
    hi_half = eng.lev_sweep("discrete", f, 100.0, n_investors=n // 2, investor_offset=n // 2, **kw)["data_T"]
    assert torch.equal(full, torch.cat([lo_half, hi_half], dim=1))


def test_grid_larger_than_one_tile(eng):
    """G = 45 > 32 runs as two grid tiles; G = 64 is the limit."""
    rs = np.random.RandomState(3)
    oc = rs.randint(0, 3, size=(700, 90)).astype(np.uint8)
    f = (1 + rs.uniform(-0.4, 0.5, size=(45, 3))).astype(np.float32)
    got = eng.lev_sweep("discrete", f, 100.0, outcomes=eng.encode_codes(oc), mode="chain")["data_T"]
    assert np.array_equal(bits(got.cpu().numpy()), bits(lo.chain_discrete(oc, f, 100.0)))


@pytest.mark.parametrize("k", [2, 3, 4])
@pytest.mark.parametrize("g", [1, 4, 7, 10, 20, 32, 33])
def test_chain_variants_agree_bit_for_bit_on_random_tables(eng, g, k):
    """FSEL, shared-memory and FMA selection are the same fp32 chain (variant 3 falls back for K = 4)."""
    rs = np.random.RandomState(100 * g + k)
    n, h = 777, 261                                   # ragged: partial row tile, partial step tile
    oc = rs.randint(0, k, size=(n, h)).astype(np.uint8)
    f = (1 + rs.uniform(-0.95, 1.5, size=(g, k))).astype(np.float32)
    want = lo.chain_discrete(oc, f, 100.0)
    codes = eng.encode_codes(oc)
    for v in (1, 2, 3):
        got = eng.lev_sweep("discrete", f, 100.0, outcomes=codes, mode="chain", variant=v)["data_T"]
        assert np.array_equal(bits(got.cpu().numpy()), bits(want)), (g, k, v)


def test_fma_variant_falls_back_when_a_delta_is_not_reproducible(eng):
    """
    Tables on which a single fused multiply-add cannot land on the factor (huge
    spread, zero, overflowing scaled delta, inf) still give the exact chain.
    """
    rs = np.random.RandomState(9)
    oc = rs.randint(0, 3, size=(300, 40)).astype(np.uint8)
    codes = eng.encode_codes(oc)
    tables = [
        np.float32([[1e-8, 0.5, 3.0], [0.0, 1.0, 2.0], [1.5, 0.0, 0.0]]),
        np.float32([[1.0, 0.5, 40.0], [1.0, 2.0, 3.0], [0.25, 0.5, 0.75]]),      # delta 39 * 2^123 overflows
        np.float32([[1.0, 0.5, np.inf], [1.0, 2.0, 3.0], [0.25, 0.5, 0.75]]),
        np.float32([[3.0000002, 0.5, 1e-8 + 1e-15], [1.0, 2.0, 3.0], [0.25, 0.5, 0.75]]),
    ]
    for f in tables:
        with np.errstate(all="ignore"):
            want = lo.chain_discrete(oc, f, 1.0)
        got = eng.lev_sweep("discrete", f, 1.0, outcomes=codes, mode="chain", variant=3)["data_T"]
        assert np.array_equal(bits(got.cpu().numpy()), bits(want))
    f2 = np.float32([[1.0, 40.0], [0.5, 0.7]])
    oc2 = (oc & 1).astype(np.uint8)
    got = eng.lev_sweep("discrete", f2, 1.0, outcomes=eng.encode_codes(oc2), mode="chain", variant=3)["data_T"]
    with np.errstate(all="ignore"):
        assert np.array_equal(bits(got.cpu().numpy()), bits(lo.chain_discrete(oc2, f2, 1.0)))


def test_empty_and_invalid_inputs(eng):
    from rlmd_b200._lib import B200Error
    f = np.ones((3, 2), np.float32)
    empty = torch.zeros((0, 16), dtype=torch.uint8, device="cuda")
    res = eng.lev_sweep("discrete", f, 100.0, outcomes=empty, mode="chain")
    assert res["data_T"].shape == (3, 0)
    with pytest.raises(B200Error):
        eng.lev_sweep("discrete", np.ones((65, 2), np.float32), 100.0, outcomes=torch.zeros((4, 16), dtype=torch.uint8, device="cuda"))
    with pytest.raises(B200Error):
        eng.lev_sweep("gbm", np.ones(3, np.float32), 100.0, outcomes=torch.zeros((4, 16), device="cuda"), mode="chain")


# ------------------------------------------------------------------ rowstats
def _rowstats_case(rs, n, kind):
    if kind == "lognormal":
        return rs.lognormal(3, 2, size=n).astype(np.float32)
    if kind == "ties":          # few distinct values: heavy ties across the thresholds
        return rs.choice(np.float32([1.5, 0.5, 2.25, 100.0]), size=n).astype(np.float32)
    if kind == "signed":
        return rs.standard_normal(n).astype(np.float32) * np.float32(3)
    if kind == "constant":
        return np.full(n, np.float32(7.25))
    if kind == "denormal":
        return (rs.lognormal(0, 1, size=n) * 1e-41).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["lognormal", "ties", "signed", "constant", "denormal"])
@pytest.mark.parametrize("n,top", [(2, 1), (5, 2), (1000, 1), (4097, 37), (100003, 10)])
def test_rowstats_vs_oracle(eng, kind, n, top):
    rs = np.random.RandomState(n + top)
    rows = 3
    v = np.stack([_rowstats_case(rs, n, kind) for _ in range(rows)])
    got = eng.rowstats(torch.from_numpy(v).cuda(), top).cpu().numpy()
    want = np.stack([lo.summary_stats(v[r], top) for r in range(rows)])
    # medians: exact
    assert np.array_equal(got[:, 9:12], want[:, 9:12])
    scale = np.abs(want[:, 0:3]).max(axis=1, keepdims=True) + 1e-300
    assert np.all(np.abs(got[:, :9] - want[:, :9]) <= 1e-11 * np.maximum(np.abs(want[:, :9]), scale))


def test_rowstats_inf_and_nan_follow_torch(eng):
    v = np.float32([1.0, np.inf, 3.0, 2.0, np.inf, 0.5, 4.0])
    got = eng.rowstats(torch.from_numpy(v[None]).cuda(), 2).cpu().numpy()[0]
    t = torch.from_numpy(v)
    s = t.sort(descending=True)[0]
    for j, grp in enumerate((t, s[:2], s[2:])):
        std, mean = torch.std_mean(grp, unbiased=False)
        for idx, w in ((0 + j, mean), (6 + j, std), (9 + j, grp.median()), (3 + j, (grp - mean).abs().mean())):
            w = float(w)
            assert (np.isnan(got[idx]) and np.isnan(w)) or got[idx] == pytest.approx(w, rel=1e-6), (idx, got[idx], w)


def _torch_stats(v: np.ndarray, top: int) -> np.ndarray:
    """The reference block (lev/lev_exp.py:177-192) itself, on torch CPU."""
    t = torch.from_numpy(v)
    s = t.sort(descending=True)[0]
    out = np.zeros(12)
    for j, grp in enumerate((t, s[:top], s[top:])):
        std, mean = torch.std_mean(grp, unbiased=False)
        out[0 + j], out[6 + j] = float(mean), float(std)
        out[3 + j] = float((grp - mean).abs().mean())
        out[9 + j] = float(grp.median())
    return out


@pytest.mark.parametrize("top", [1, 3, 8, 15])
def test_rowstats_rows_in_every_key_mode_in_one_launch(eng, top):
    """
    Passes 1..3 pick a key mode per row from pass 0's histogram: raw bits for an
    all-positive row, the sign flip for a row with the sign bit set somewhere
    (-0.0 included), the full map for a row with NaNs (either sign bit).
    """
    rs = np.random.RandomState(top)
    n = 29
    rows = [rs.lognormal(0, 2, n), rs.standard_normal(n), rs.lognormal(0, 1, n), rs.lognormal(0, 1, n),
            rs.standard_normal(n), np.abs(rs.standard_normal(n))]
    v = np.stack(rows).astype(np.float32)
    v[2, 5] = np.float32(-0.0)                                   # sign bit, value zero
    v[3, [1, 7]] = np.float32(np.nan)
    v[4, [0, 9, 20]] = np.array([0xFFC00000, 0x7FC00000, 0xFF800001], dtype=np.uint32).view(np.float32)  # -NaN, NaN, -sNaN
    v[5, [2, 3]] = np.float32(np.inf)
    got = eng.rowstats(torch.from_numpy(v).cuda(), top).cpu().numpy()
    for r in range(v.shape[0]):
        want = _torch_stats(v[r], top)
        for idx in range(12):
            assert (np.isnan(got[r, idx]) and np.isnan(want[idx])) or got[r, idx] == pytest.approx(want[idx], rel=2e-6, abs=1e-7), \
                (r, idx, got[r], want)
        ok = ~np.isnan(want[9:12])
        assert np.array_equal(got[r, 9:12][ok], want[9:12][ok]), "order statistics are exact"


def test_rowstats_large_top_group(eng):
    """top ~ n/2: the 'above thr' side is no longer rare (it sits behind the branch pass 2 takes seldom)."""
    rs = np.random.RandomState(3)
    v = rs.lognormal(1, 1.5, size=(2, 50_001)).astype(np.float32)
    for top in (25_000, 50_000):
        got = eng.rowstats(torch.from_numpy(v).cuda(), top).cpu().numpy()
        want = np.stack([lo.summary_stats(v[r], top) for r in range(2)])
        assert np.array_equal(got[:, 9:12], want[:, 9:12])
        assert np.allclose(got[:, :9], want[:, :9], rtol=1e-11)


def test_rowstats_strided_rows(eng):
    rs = np.random.RandomState(0)
    buf = torch.from_numpy(rs.lognormal(size=(4, 1000)).astype(np.float32)).cuda()
    view = buf[:, 3:900]
    got = eng.rowstats(view, 5).cpu().numpy()
    want = np.stack([lo.summary_stats(view[r].cpu().numpy(), 5) for r in range(4)])
    assert np.array_equal(got[:, 9:12], want[:, 9:12])
    assert np.allclose(got[:, :9], want[:, :9], rtol=1e-11)
