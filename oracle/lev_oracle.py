"""
TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy) of the reference's
optimal-leverage sweeps.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product
(rlmd_b200/) never does.

Parity pin: the reference holds no golden vectors for this path (SURVEY.md
section 8c).  This restatement is pinned instead against the reference's own
functions run in the build container on identical pre-drawn outcome arrays:
`tests/golden/gen_golden.py` wrote `tests/golden/lev_*.npz` from the unmodified
`/root/reference/lev/lev_exp.py`, and `tests/test_oracle_lev.py` checks this
file against those fixtures (bit-exact on data_T and medians).

What is restated (reference file:line):
  param_range               lev/lev_exp.py:29-53
  leverage grid + sign flip lev/lev_exp.py:80-81,160-161,534-535,620-621,1153-1154,1249-1250
  coin factors              lev/lev_exp.py:85,167-168
  dice factors              lev/lev_exp.py:541-543,631-633
  dice_sh factors           lev/lev_exp.py:1160-1166,1260-1266
  gbm factors               lev/lev_exp.py:965,1050
  sequential chain          lev/lev_exp.py:170-175 (and the three siblings)
  summary statistic block   lev/lev_exp.py:177-192 (inlined 13x in the reference)
  output row order          lev/lev_exp.py:194-210
  optimal leverage          lev/lev_exp.py:240-267, 704-738
  big-brain recurrence      lev/lev_exp.py:310-450, 784-930

All wealth arithmetic is IEEE fp32 exactly as torch-CPU performs it (NumPy
float32 arrays, one rounding per operation); the statistics are accumulated in
fp64 and are compared with a tolerance (torch reduces in fp32 with an
implementation-defined order).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

STAT_ROWS = (
    "mean", "mean_top", "mean_adj",
    "mad", "mad_top", "mad_adj",
    "std", "std_top", "std_adj",
    "med", "med_top", "med_adj",
)


# --------------------------------------------------------------------------- grid
def param_range(low: float, high: float, increment: float) -> list:
    """Leverage grid as Python doubles, truncation quirks included (App. A)."""
    first = int(low / increment)
    stop = int(high / increment + 1)
    frac = low / increment - first
    grid = [(k + frac) * increment for k in range(first, stop)]
    if len(grid) > 1 and 0 in grid:
        grid.remove(0)
    return grid


def lev_grid(low: float, high: float, incr: float, up_r=None, down_r=None) -> np.ndarray:
    """fp32 grid; negated iff -down_r > up_r (never for GBM: pass None)."""
    g = np.asarray(param_range(low, high, incr), dtype=F32)
    if up_r is not None and -down_r > up_r:
        g = -g
    return g


# ------------------------------------------------------------------------ factors
def coin_factors(lev: np.ndarray, up_r: float, down_r: float) -> np.ndarray:
    """[G,2] fp32, column = outcome value (0 -> down, 1 -> up)."""
    lev = lev.astype(F32)
    one = F32(1)
    up = one + lev * F32(up_r)
    dn = one + lev * F32(down_r)
    return np.stack([dn, up], axis=1).astype(F32)


def dice_factors(lev: np.ndarray, up_r: float, down_r: float, mid_r: float) -> np.ndarray:
    """[G,3] fp32, column = outcome (0 up, 1 down, 2 mid)."""
    lev = lev.astype(F32)
    one = F32(1)
    cols = [one + lev * F32(r) for r in (up_r, down_r, mid_r)]
    return np.stack(cols, axis=1).astype(F32)


def dice_sh_factors(lev, up_r, down_r, mid_r, sh_up_r, sh_down_r, sh_mid_r) -> np.ndarray:
    """[G,3] fp32: (1 + l*r) + (1-l)*sh, evaluated left to right in fp32."""
    lev = lev.astype(F32)
    one = F32(1)
    cols = []
    for r, sh in ((up_r, sh_up_r), (down_r, sh_down_r), (mid_r, sh_mid_r)):
        cols.append((one + lev * F32(r)) + (one - lev) * F32(sh))
    return np.stack(cols, axis=1).astype(F32)


def general_factors(a: np.ndarray, b: np.ndarray, r, sh) -> np.ndarray:
    """Engine-only 2-D grid: m[g][k] = (1 + a_g r_k) + b_g sh_k (fp32)."""
    a = a.astype(F32)
    b = b.astype(F32)
    one = F32(1)
    cols = [(one + a * F32(rk)) + b * F32(sk) for rk, sk in zip(r, sh)]
    return np.stack(cols, axis=1).astype(F32)


# -------------------------------------------------------------------------- chain
def chain_discrete(outcomes: np.ndarray, factors: np.ndarray, value_0: float,
                   keep_steps: bool = False):
    """
    Sequential fp32 product ((V0*m_0)*m_1)*... per (grid point, investor).

    outcomes: [N,H] integer codes; factors: [G,K] fp32.
    Returns data_T [G,N] fp32 (wealth after the last step) and, when
    keep_steps, the list over t = 1..H-1 of [G,N] wealth arrays.
    """
    oc = np.asarray(outcomes)
    n, h = oc.shape
    g = factors.shape[0]
    w = np.full((g, n), F32(value_0), dtype=F32)
    steps = []
    for t in range(h):
        m = factors[:, oc[:, t].astype(np.int64)]  # [G,N] fp32
        w = (w * m).astype(F32)
        if keep_steps and t >= 1:
            steps.append(w.copy())
    return (w, steps) if keep_steps else w


def chain_gbm(x: np.ndarray, lev: np.ndarray, value_0: float, keep_steps: bool = False):
    """GBM: factor = expf(fl32(l*x)); then the same sequential fp32 chain."""
    x = np.asarray(x, dtype=F32)
    n, h = x.shape
    lev = lev.astype(F32)
    w = np.full((lev.shape[0], n), F32(value_0), dtype=F32)
    steps = []
    with np.errstate(over="ignore", invalid="ignore"):
        for t in range(h):
            m = np.exp((lev[:, None] * x[None, :, t]).astype(F32)).astype(F32)
            w = (w * m).astype(F32)
            if keep_steps and t >= 1:
                steps.append(w.copy())
    return (w, steps) if keep_steps else w


# -------------------------------------------------------- log-domain (engine adds)
def counts_discrete(outcomes: np.ndarray, k: int) -> np.ndarray:
    """[N,K] int64: how often each outcome code occurs per investor."""
    oc = np.asarray(outcomes).astype(np.int64)
    return np.stack([(oc == j).sum(axis=1) for j in range(k)], axis=1)


def log_wealth_discrete(outcomes: np.ndarray, factors: np.ndarray, value_0: float) -> np.ndarray:
    """[G,N] fp64: log V0 + sum_k n_k log m_k (fp32-rounded factors, fp64 logs)."""
    cnt = counts_discrete(outcomes, factors.shape[1]).astype(np.float64)  # [N,K]
    with np.errstate(divide="ignore", invalid="ignore"):
        lm = np.log(factors.astype(np.float64))  # [G,K]
    out = np.full((factors.shape[0], cnt.shape[0]), math.log(float(F32(value_0))))
    for j in range(factors.shape[1]):
        # a count of zero must not turn log(0) = -inf into nan
        term = np.where(cnt[None, :, j] > 0, cnt[None, :, j] * lm[:, j:j + 1], 0.0)
        out = out + term
    return out


def log_wealth_gbm(x: np.ndarray, lev: np.ndarray, value_0: float) -> np.ndarray:
    """[G,N] fp64: log V0 + sum_t fl32(l*x_t) accumulated in fp64."""
    x = np.asarray(x, dtype=F32)
    lev = lev.astype(F32)
    out = np.empty((lev.shape[0], x.shape[0]))
    for gi, l in enumerate(lev):
        out[gi] = (l * x).astype(F32).astype(np.float64).sum(axis=1)
    return out + math.log(float(F32(value_0)))


def valid_count(data_T: np.ndarray) -> np.ndarray:
    """Investors whose reference-dtype (fp32) wealth is finite and > 0."""
    return (np.isfinite(data_T) & (data_T > 0)).sum(axis=-1).astype(np.int64)


# ------------------------------------------------- packed outcomes (engine format)
def pack_codes(codes: np.ndarray, ld: int = 0) -> np.ndarray:
    """
    [N,H] codes < 4 -> uint8 [N, ld]: four 2-bit codes per byte, step t in byte
    t >> 2 at bits 2*(t&3)..+1, pad bits zero (include/rlmd_b200.h, outcome_bits = 2).
    The reference has no such format; this restates the engine's layout so that
    the packed sweep can be checked against the reference-pinned counts.
    """
    c = np.asarray(codes).astype(np.uint8) & 3
    n, h = c.shape
    nb = (h + 3) // 4
    ld = max(int(ld), nb)
    pad = np.zeros((n, nb * 4), dtype=np.uint8)
    pad[:, :h] = c
    q = pad.reshape(n, nb, 4)
    out = np.zeros((n, ld), dtype=np.uint8)
    out[:, :nb] = q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)
    return out


def unpack_codes(packed: np.ndarray, horizon: int) -> np.ndarray:
    p = np.asarray(packed, dtype=np.uint8)
    c = np.stack([(p >> s) & 3 for s in (0, 2, 4, 6)], axis=-1)
    return c.reshape(p.shape[0], -1)[:, :horizon]


# --------------------------------------------------------------------- statistics
def lower_median(v: np.ndarray) -> float:
    """torch.median: ascending order statistic at 0-based index (n-1)//2."""
    s = np.sort(v)
    return s[(s.shape[0] - 1) // 2]


def group_stats(v: np.ndarray):
    """(mean, mad, std, med) of one group, fp64 accumulation, population std."""
    v64 = v.astype(np.float64)
    med = float("nan") if np.isnan(v).any() else float(lower_median(v))
    if not np.isfinite(v64).all():
        # torch.std_mean runs Welford's update: one inf (then inf - inf) makes the
        # mean, hence std and MAD, nan (seen in tests/golden/lev_coin_overflow.npz)
        # (a one-element group keeps its element as the mean: the first update is exact)
        mean = float(v64[0]) if v64.shape[0] == 1 else float("nan")
        return mean, float("nan"), float("nan"), med
    with np.errstate(over="ignore", invalid="ignore"):
        mean = v64.mean()
        mad = np.abs(v64 - mean).mean()
        std = math.sqrt(((v64 - mean) ** 2).mean())
    return mean, mad, std, med


def summary_stats(v: np.ndarray, top: int) -> np.ndarray:
    """12 statistics of one wealth vector in the reference's row order."""
    s = np.sort(v)[::-1]
    groups = (v, s[:top], s[top:])
    res = [group_stats(x) for x in groups]
    out = np.empty(12)
    for gi in range(3):
        mean, mad, std, med = res[gi]
        out[0 + gi], out[3 + gi], out[6 + gi], out[9 + gi] = mean, mad, std, med
    return out


def percentile_type8(v: np.ndarray, q: float) -> float:
    """Hyndman-Fan type 8 (numpy method="median_unbiased"), App. B."""
    s = np.sort(v.astype(np.float64))
    n = s.shape[0]
    virt = n * q + (1 + q) / 3 - 1
    virt = min(max(virt, 0.0), n - 1.0)
    lo = int(math.floor(virt))
    hi = min(lo + 1, n - 1)
    return float(s[lo] + (virt - lo) * (s[hi] - s[lo]))


# ------------------------------------------------------- reference entry points
def smart_lev(outcomes, factors_or_lev, lev, top, value_0, gbm=False):
    """
    Restates *_smart_lev: data [G,13,H-1] fp32, data_T [G,N] fp32.
    `lev` is the fp32 grid stored in row 12.
    """
    if gbm:
        data_T, steps = chain_gbm(outcomes, factors_or_lev, value_0, keep_steps=True)
    else:
        data_T, steps = chain_discrete(outcomes, factors_or_lev, value_0, keep_steps=True)
    g = data_T.shape[0]
    data = np.zeros((g, 13, len(steps)), dtype=F32)
    with np.errstate(over="ignore", invalid="ignore"):
        for t, w in enumerate(steps):
            for gi in range(g):
                data[gi, :12, t] = summary_stats(w[gi], top).astype(F32)
                data[gi, 12, t] = lev[gi]
    return data, data_T


def fixed_final_lev(outcomes, factors_or_lev, top, value_0, gbm=False) -> np.ndarray:
    """
    Restates *_fixed_final_lev up to the product order: returns the [G,12]
    statistics of the sequential-chain wealth (the reference only prints them,
    with three significant digits, from a torch.prod of unspecified order).
    """
    w = chain_gbm(outcomes, factors_or_lev, value_0) if gbm else \
        chain_discrete(outcomes, factors_or_lev, value_0)
    with np.errstate(over="ignore", invalid="ignore"):
        return np.stack([summary_stats(w[gi], top) for gi in range(w.shape[0])])


def format_final(lev: np.ndarray, stats: np.ndarray) -> str:
    """Text of the reference's per-leverage print (lev/lev_exp.py:106-125)."""
    lines = []
    for l, s in zip(lev, stats):
        mean, mean_top, mean_adj, mad, mad_top, mad_adj, std, std_top, std_adj, med, med_top, med_adj = s
        lines.append(
            """       lev {:1.0f}%:
                 avg mean/med/mad/std:  $ {:1.2e} / {:1.2e} / {:1.1e} / {:1.1e}
                 top mean/med/mad/std:  $ {:1.2e} / {:1.2e} / {:1.1e} / {:1.1e}
                 adj mean/med/mad/std:  $ {:1.2e} / {:1.2e} / {:1.1e} / {:1.1e}""".format(
                float(l) * 100, mean, med, mad, std, mean_top, med_top, mad_top, std_top,
                mean_adj, med_adj, mad_adj, std_adj))
    return "\n".join(lines)


# ---------------------------------------------------------------------- big brain
def _opt_lev_f32(v, v0, vmin, eta32, roll):
    """
    coin_optimal_lev / the retention branch of dice_optimal_lev (lev/lev_exp.py:258-267,
    :730-738) on an fp32 vector: every operand fp32 (the float64 0-dim LEV_FACTOR of
    the scripts does not promote an fp32 [N] tensor, it is rounded to fp32).
    """
    one = F32(1)
    if roll == 0:
        return (eta32 * (one - vmin / v)).astype(F32)
    floor = np.where(v <= v0, vmin, v0 + roll * (v - v0)).astype(F32)
    return (eta32 * (one - floor / v)).astype(F32)


def _initial_lev_f64(v0, vmin, eta64) -> float:
    """
    The leverage of step 0 (:327, :806-808): every operand is 0-dim, so the fp32 inner
    term times the float64 LEV_FACTOR promotes to float64 (both branches of the rule
    reduce to value_min because value_t == value_0).
    """
    inner = F32(1) - F32(vmin / v0)
    return float(eta64) * float(F32(inner))


def big_brain(kind, outcomes, top, value_0, returns, lev_factor, stop_grid, roll_grid, want_wealth=False):
    """
    coin_big_brain_lev (lev/lev_exp.py:270-452) / dice_big_brain_lev (:741-932):
    data [R,S,26,H-1] fp32 - rows 0..11 wealth statistics after the step, 12..23
    leverage statistics before it, 24 stop-loss, 25 retention ratio.

    kind "coin": outcomes {0,1}, returns (up, down); the whole chain is fp32
        g = fl32(return); V <- V * (1 + lev * g); lev <- opt(V).
    kind "dice": outcomes {0,1,2}, returns (up, down, mid); the reference casts the
        outcomes to float64 (:791) so returns and the WEALTH CHAIN are float64; with
        retention 0 the leverage is float64 too (:727-728), otherwise it is computed
        in fp32 from the fp32-rounded wealth (:731-738) and promoted in the update.
    lev_factor: the float64 value of the scripts' LEV_FACTOR tensor.
    """
    oc = np.asarray(outcomes)
    n, h = oc.shape
    stop_grid = np.asarray(stop_grid, dtype=F32)
    roll_grid = np.asarray(roll_grid, dtype=F32)
    v0 = F32(value_0)
    eta64, eta32 = float(lev_factor), F32(lev_factor)
    if kind == "coin":
        ret = np.where(oc == 1, F32(returns[0]), F32(returns[1])).astype(F32)
    else:
        ret = np.where(oc == 0, float(returns[0]), np.where(oc == 1, float(returns[1]), float(returns[2])))
    data = np.zeros((len(roll_grid), len(stop_grid), 26, max(h - 1, 0)), dtype=F32)
    wealth = np.zeros((len(roll_grid), len(stop_grid), n), dtype=np.float64)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for j, roll in enumerate(roll_grid):
            for i, stop in enumerate(stop_grid):
                vmin = F32(stop * v0)
                lev0 = _initial_lev_f64(v0, vmin, eta64)
                if kind == "coin":
                    v = (v0 * (F32(1) + F32(lev0) * ret[:, 0])).astype(F32)
                    lev = _opt_lev_f32(v, v0, vmin, eta32, roll)
                else:
                    v = float(v0) * (1.0 + lev0 * ret[:, 0])
                    lev = (eta64 * (1.0 - float(vmin) / v)) if roll == 0 else \
                        _opt_lev_f32(v.astype(F32), v0, vmin, eta32, roll)
                for t in range(h - 1):
                    data[j, i, 12:24, t] = summary_stats(lev, top).astype(F32)
                    data[j, i, 24, t] = stop
                    data[j, i, 25, t] = roll
                    if kind == "coin":
                        v = (v * (F32(1) + lev * ret[:, t + 1])).astype(F32)
                        lev = _opt_lev_f32(v, v0, vmin, eta32, roll)
                    else:
                        v = v * (1.0 + lev.astype(np.float64) * ret[:, t + 1])
                        lev = (eta64 * (1.0 - float(vmin) / v)) if roll == 0 else \
                            _opt_lev_f32(v.astype(F32), v0, vmin, eta32, roll)
                    data[j, i, 0:12, t] = summary_stats(v, top).astype(F32)
                wealth[j, i] = v
    return (data, wealth) if want_wealth else data


def galaxy_brain(ru_grid, rd_grid, pu_grid) -> np.ndarray:
    """coin_galaxy_brain_lev (lev/lev_exp.py:455-505): Kelly table, fp32."""
    data = np.zeros((len(pu_grid), len(ru_grid), len(ru_grid), 4), dtype=F32)
    for i, pu in enumerate(pu_grid):
        for j, ru in enumerate(ru_grid):
            for k, rd in enumerate(rd_grid):
                data[i, j, k] = (pu, ru, rd, pu / rd - (1 - pu) / ru)
    return data


def growth_summary(log_w: np.ndarray, data_T, value_0: float, horizon: int, quantiles=(0.05, 0.5)) -> np.ndarray:
    """
    Time-average growth rates g = (log W_T - log V0) / H per leverage row, reduced the
    way the reference's consumers reduce per-run quantities: mean / std with NumPy,
    percentiles with np.percentile(method="median_unbiased")
    (tools/eval_episodes.py:276-315, plotting/plots_multiverse.py:167-169).
    Columns: valid runs (fp32 wealth finite and > 0), mean, population std, mean over the
    valid runs, min, max, then the quantiles.
    """
    lw = np.asarray(log_w, dtype=np.float64)
    out = np.zeros((lw.shape[0], 6 + len(quantiles)))
    for r in range(lw.shape[0]):
        g = (lw[r] - np.log(float(value_0))) / float(horizon)
        if data_T is not None:
            w = np.asarray(data_T[r], dtype=np.float32)
            ok = np.isfinite(w) & (w > 0)
        else:
            ok = np.isfinite(lw[r])
        with np.errstate(invalid="ignore"):
            out[r, 0] = ok.sum()
            out[r, 1] = g.mean()
            out[r, 2] = np.sqrt(np.mean((g - g.mean()) ** 2))
            out[r, 3] = g[ok].mean() if ok.any() else np.nan
            out[r, 4], out[r, 5] = g.min(), g.max()
            for j, q in enumerate(quantiles):
                out[r, 6 + j] = np.percentile(g, 100.0 * q, method="median_unbiased")
    return out
