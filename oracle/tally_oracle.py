"""
TEST INFRASTRUCTURE ONLY (never imported by rlmd_b200/): a NumPy restatement of the
count-tuple route the GPU takes for the discrete *_fixed_final_lev sweeps
(rlmd_b200/csrc/tally.cu), so that the ALGORITHM - "the 12 statistics of every
leverage follow from the distinct outcome-count tuples and their multiplicities" -
is pinned on the CPU against the per-investor statistics of oracle/lev_oracle.py,
which are pinned to the reference's own output (lev/lev_exp.py:89-104, :551-562;
tests/golden/lev_*.npz, tests/golden/final_text.json).

  bins      : distinct tuples (n_0 .. n_{K-1}) with their investor counts
  wealth    : fl32(exp(log V0 + sum_k n_k log m_k)) per bin and leverage - the LOG
              sweep's data_T value of every investor in the bin
  order statistics : weighted lower medians (ascending index (n-1)//2) of all /
              top-K / the rest; the value at descending rank K is the threshold,
              its ties are apportioned: K - #(w > thr) of them belong to the top
  moments   : fp64, two-pass, weighted; the top / rest sums are taken over
              {w > thr}, {w < thr} plus the tie shares (never all-minus-top)
"""
import math

import numpy as np

from . import lev_oracle as lo

F32 = np.float32


def bins_of(outcomes: np.ndarray, n_outcomes: int):
    """(tuples [B,K] int64, counts [B] int64) in first-occurrence order."""
    cnt = lo.counts_discrete(outcomes, n_outcomes).astype(np.int64)
    tuples, first, mult = np.unique(cnt, axis=0, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    return tuples[order], mult[order].astype(np.int64)


def bin_wealth(tuples: np.ndarray, factors: np.ndarray, value_0: float) -> np.ndarray:
    """[G,B] fp32: the data_T value shared by the investors of a bin (same expression as log_wealth_discrete)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        lm = np.log(factors.astype(np.float64))
    lw = np.full((factors.shape[0], tuples.shape[0]), math.log(float(F32(value_0))))
    for k in range(factors.shape[1]):
        nk = tuples[None, :, k].astype(np.float64)
        with np.errstate(invalid="ignore"):                       # 0 * -inf in the branch np.where discards
            lw = lw + np.where(nk > 0, nk * lm[:, k:k + 1], 0.0)
    with np.errstate(over="ignore", invalid="ignore"):
        return np.exp(lw).astype(F32)


def _weighted_lower_median(vals_asc: np.ndarray, w_asc: np.ndarray) -> float:
    n = int(w_asc.sum())
    pos = (n - 1) // 2
    return float(vals_asc[np.searchsorted(np.cumsum(w_asc), pos, side="right")])


def _moments(vals: np.ndarray, w: np.ndarray):
    """(mean, mad, std) of the multiset, torch semantics for non-finite members (lev_oracle.group_stats)."""
    n = int(w.sum())
    v64 = vals.astype(np.float64)
    if not np.isfinite(v64[w > 0]).all():
        return (float(v64[w > 0][0]) if n == 1 else float("nan")), float("nan"), float("nan")
    wf = w.astype(np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        mean = float((wf * v64).sum() / n)
        mad = float((wf * np.abs(v64 - mean)).sum() / n)
        std = math.sqrt(float((wf * (v64 - mean) ** 2).sum() / n))
    return mean, mad, std


def stats_from_bins(wealth: np.ndarray, counts: np.ndarray, top: int) -> np.ndarray:
    """[12] in the reference's row order from one leverage's bin wealth [B] and the bin counts [B]."""
    n = int(counts.sum())
    k = int(top)
    has_nan = bool(np.isnan(wealth[counts > 0]).any())
    order = np.argsort(-wealth.astype(np.float64), kind="stable")      # descending, NaN last for argsort of -x
    if has_nan:                                                        # torch.sort puts NaN first in descending order
        nan = np.isnan(wealth[order])
        order = np.concatenate([order[nan], order[~nan]])
    v, c = wealth[order], counts[order]
    cum = np.cumsum(c)
    j = int(np.searchsorted(cum, k - 1, side="right"))                 # the bin that holds descending rank K-1
    in_top = c.copy()
    in_top[j + 1:] = 0
    in_top[j] = k - (int(cum[j - 1]) if j > 0 else 0)                   # the tie share of the threshold bin(s)
    in_adj = c - in_top
    out = np.empty(12)
    groups = ((v, c), (v, in_top), (v, in_adj))
    for gi, (vals, w) in enumerate(groups):
        keep = w > 0
        vals_g, w_g = vals[keep], w[keep]
        mean, mad, std = _moments(vals_g, w_g)
        asc = np.argsort(vals_g.astype(np.float64), kind="stable")
        med = float("nan") if np.isnan(vals_g).any() else _weighted_lower_median(vals_g[asc], w_g[asc])
        out[0 + gi], out[3 + gi], out[6 + gi], out[9 + gi] = mean, mad, std, med
    assert int(in_top.sum()) == k and int(in_adj.sum()) == n - k
    return out


def final_stats(outcomes: np.ndarray, factors: np.ndarray, top: int, value_0: float) -> np.ndarray:
    """[G,12]: what b200_lev_tally -> b200_tally_finalize -> b200_tally_stats produce."""
    tuples, counts = bins_of(outcomes, factors.shape[1])
    w = bin_wealth(tuples, factors, value_0)
    return np.stack([stats_from_bins(w[g], counts, top) for g in range(w.shape[0])])
