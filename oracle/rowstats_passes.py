"""
TEST INFRASTRUCTURE ONLY - NumPy restatement of the PASS STRUCTURE of
rlmd_b200/csrc/rowstats.cu (the four streaming passes + the resolve steps of the
3-level radix select), writing its partial sums / histograms at the same
workspace words the CUDA kernels use (b200_rowstats_exchange).

Purpose: the multi-GPU exchange protocol (rlmd_b200/sharding.py:exchange_phases)
can be run on CPU under gloo with this class standing in for the kernels, and
its result checked against the sort-based statistics of the reference
(oracle/lev_oracle.py:summary_stats, lev/lev_exp.py:177-192).
"""
import numpy as np

L1, L2, L3, NT = 11, 11, 10, 4


def float_key(x: np.ndarray) -> np.ndarray:
    """Order-preserving fp32 -> uint32 map in torch.sort order (NaN greatest)."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    neg = (b >> np.uint64(31)) == 1
    k = np.where(neg, (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    k = np.where((b & np.uint64(0x7FFFFFFF)) > np.uint64(0x7F800000), np.uint64(0xFFFFFFFF), k)
    return k.astype(np.uint64)


def key_float(k: int) -> float:
    if k == 0xFFFFFFFF:
        return float("nan")
    b = (k & 0x7FFFFFFF) if (k & 0x80000000) else (~k & 0xFFFFFFFF)
    return float(np.array([b], dtype=np.uint32).view(np.float32)[0])


def _find_bin(hist: np.ndarray, rank: int):
    cum = np.cumsum(hist)
    b = int(np.searchsorted(cum, rank, side="right"))
    b = min(b, len(hist) - 1)
    return b, int(rank - (cum[b - 1] if b > 0 else 0))


class RowStatsPasses:
    """run(phase) for phases 0..4 on `values` [rows, n_local]; `ws` is the int64 [rows, words] workspace."""

    def __init__(self, values: np.ndarray, n_total: int, top: int, ws: np.ndarray, offsets: dict):
        self.v = np.ascontiguousarray(values, dtype=np.float32)
        self.keys = float_key(self.v)
        self.n, self.K = int(n_total), int(top)
        self.ws = ws
        self.wsf = ws.view(np.float64)
        self.off = offsets          # {"h1","h2","h3","cnt"}: int64 word offsets inside a row
        rows = self.v.shape[0]
        self.prefix = np.zeros((rows, NT), dtype=np.uint64)
        self.rank = np.zeros((rows, NT), dtype=np.int64)
        self.value = np.zeros((rows, NT))
        self.mean = np.zeros((rows, 3))
        self.ties = np.zeros((rows, 2), dtype=np.int64)
        self.nonfinite = np.zeros((rows, 3), dtype=bool)
        self.has_nan = np.zeros((rows, 3), dtype=bool)
        self.stats = np.zeros((rows, 12))

    def run(self, phase: int) -> None:
        n, K, off = self.n, self.K, self.off
        for r in range(self.v.shape[0]):
            x, k = self.v[r].astype(np.float64), self.keys[r]
            I, D = self.ws[r], self.wsf[r]
            if phase == 0:
                I[:] = 0
                with np.errstate(invalid="ignore"):
                    D[0] = x.sum()
                I[off["h1"]:off["h1"] + (1 << L1)] = np.bincount((k >> np.uint64(32 - L1)).astype(np.int64),
                                                                 minlength=1 << L1)
            elif phase == 1:
                h1 = I[off["h1"]:off["h1"] + (1 << L1)]
                for j, rk in enumerate(((n - 1) // 2, n - K, n - K + (K - 1) // 2, (n - K - 1) // 2)):
                    b, rem = _find_bin(h1, rk)
                    self.prefix[r, j], self.rank[r, j] = b << (32 - L1), rem
                self.mean[r, 0] = D[0] / n
                ninf, pinf, nan = int(h1[3]), int(h1[2044]), int(h1[2047])
                hi = pinf + nan
                self.nonfinite[r] = [(hi + ninf) > 0, hi > 0 or ninf > n - K, hi > K or ninf > 0]
                self.has_nan[r] = [nan > 0, nan > 0, nan > K]
                top_bits = k >> np.uint64(32 - L1)
                mid = ((k >> np.uint64(L3)) & np.uint64((1 << L2) - 1)).astype(np.int64)
                for j in range(NT):
                    sel = top_bits == (self.prefix[r, j] >> np.uint64(32 - L1))
                    lo = off["h2"] + j * (1 << L2)
                    I[lo:lo + (1 << L2)] = np.bincount(mid[sel], minlength=1 << L2)
            elif phase == 2:
                for j in range(NT):
                    lo = off["h2"] + j * (1 << L2)
                    b, rem = _find_bin(I[lo:lo + (1 << L2)], int(self.rank[r, j]))
                    self.prefix[r, j] |= np.uint64(b << L3)
                    self.rank[r, j] = rem
                low = (k & np.uint64((1 << L3) - 1)).astype(np.int64)
                hi = k >> np.uint64(L3)
                for j in range(NT):
                    sel = hi == (self.prefix[r, j] >> np.uint64(L3))
                    lo = off["h3"] + j * (1 << L3)
                    I[lo:lo + (1 << L3)] = np.bincount(low[sel], minlength=1 << L3)
                # coarse split about thr's 22-bit prefix (the rest comes from thr's level-3 bins)
                thr_hi = self.prefix[r, 1] >> np.uint64(L3)
                gt, lt = hi > thr_hi, hi < thr_hi
                with np.errstate(invalid="ignore"):
                    D[1], D[2] = x[gt].sum(), x[lt].sum()
                I[off["cnt"]] = int(gt.sum())     # the count below follows from the total (phase 3)
            elif phase == 3:
                for j in range(NT):
                    lo = off["h3"] + j * (1 << L3)
                    b, rem = _find_bin(I[lo:lo + (1 << L3)], int(self.rank[r, j]))
                    self.prefix[r, j] |= np.uint64(b)
                    self.rank[r, j] = rem
                    self.value[r, j] = key_float(int(self.prefix[r, j]))
                thr = int(self.prefix[r, 1])
                thr_lo, base = thr & ((1 << L3) - 1), thr & ~((1 << L3) - 1)
                h3 = I[off["h3"] + (1 << L3):off["h3"] + 2 * (1 << L3)]
                n_gt, s_gt, s_lt = int(I[off["cnt"]]), D[1], D[2]
                n_lt = n - n_gt - int(h3.sum())
                f_gt = f_lt = 0.0
                c_gt = c_lt = 0
                with np.errstate(invalid="ignore", over="ignore"):
                    for lo in np.nonzero(h3)[0]:
                        c = int(h3[lo])
                        if lo > thr_lo:
                            f_gt, c_gt = f_gt + c * key_float(base | int(lo)), c_gt + c
                        elif lo < thr_lo:
                            f_lt, c_lt = f_lt + c * key_float(base | int(lo)), c_lt + c
                    if c_gt:
                        s_gt = s_gt + f_gt
                    if c_lt:
                        s_lt = s_lt + f_lt
                n_gt, n_lt = n_gt + c_gt, n_lt + c_lt
                thr_v = self.value[r, 1]
                n_eq = n - n_gt - n_lt
                tt = K - n_gt
                self.ties[r] = [tt, n_eq - tt]
                with np.errstate(invalid="ignore", over="ignore"):
                    self.mean[r, 1] = (s_gt + (tt * thr_v if tt > 0 else 0.0)) / K
                    self.mean[r, 2] = (s_lt + ((n_eq - tt) * thr_v if n_eq - tt > 0 else 0.0)) / (n - K)
                    gt, lt = k > np.uint64(thr), k < np.uint64(thr)
                    dt, da = x[gt] - self.mean[r, 1], x[lt] - self.mean[r, 2]
                    D[5], D[6], D[7], D[8] = np.abs(dt).sum(), (dt * dt).sum(), np.abs(da).sum(), (da * da).sum()
                    d = x - self.mean[r, 0]           # the all-group moments ride on this pass too
                    D[3], D[4] = np.abs(d).sum(), (d * d).sum()
            else:
                thr_v = self.value[r, 1]
                dt, da = thr_v - self.mean[r, 1], thr_v - self.mean[r, 2]
                tt, ta = float(self.ties[r, 0]), float(self.ties[r, 1])
                abs_top = D[5] + (tt * abs(dt) if tt > 0 else 0.0)
                sq_top = D[6] + (tt * dt * dt if tt > 0 else 0.0)
                abs_adj = D[7] + (ta * abs(da) if ta > 0 else 0.0)
                sq_adj = D[8] + (ta * da * da if ta > 0 else 0.0)
                s = self.stats[r]
                s[0:3] = self.mean[r]
                s[3:6] = [D[3] / n, abs_top / K, abs_adj / (n - K)]
                with np.errstate(invalid="ignore"):
                    s[6:9] = [np.sqrt(D[4] / n), np.sqrt(sq_top / K), np.sqrt(sq_adj / (n - K))]
                s[9:12] = [self.value[r, 0], self.value[r, 2], self.value[r, 3]]
                size = (n, K, n - K)
                single = (np.nan, self.value[r, 1], self.value[r, 3])
                for j in range(3):
                    if self.nonfinite[r, j]:
                        s[j] = single[j] if size[j] == 1 else np.nan
                        s[3 + j] = s[6 + j] = np.nan
                    if self.has_nan[r, j]:
                        s[9 + j] = np.nan
