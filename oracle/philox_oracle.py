"""
TEST INFRASTRUCTURE ONLY - NumPy restatement of Philox4x32-10 (Salmon et al.,
"Parallel random numbers: as easy as 1, 2, 3", SC'11; the same generator as
cuRAND's curand_philox4x32_x.h) and of the engine's draw conventions, so that
the outcome arrays an on-device Philox sweep consumes can be reproduced
bit-for-bit on the CPU ("bit-exact outcome indexing").

Known-answer vectors: Random123's kat_vectors for philox4x32 10 rounds
(checked in tests/test_oracle_philox.py).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
TAG_LEV, TAG_ENV, TAG_REPLAY = 0x4C455600, 0x454E5600, 0x52504C00
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def lev_words(seed: int, investor_ids: np.ndarray, horizon: int) -> np.ndarray:
    """uint32 [N,H]: word t of investor i = output (t % 4) of block t // 4."""
    ids = np.asarray(investor_ids, dtype=np.uint64)
    nblk = (horizon + 3) // 4
    j = np.arange(nblk, dtype=np.uint64)[None, :]
    out = philox4x32_10((ids & MASK)[:, None], (ids >> np.uint64(32))[:, None], j, TAG_LEV,
                        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(out, axis=-1).reshape(ids.shape[0], nblk * 4)
    return words[:, :horizon]


def thresholds(probs):
    thr, acc = [], 0.0
    for p in probs[:-1]:
        acc += float(p)
        thr.append(min(int(np.floor(acc * 4294967296.0)), 0xFFFFFFFF))
    return thr


def discrete_codes(seed, investor_ids, horizon, probs) -> np.ndarray:
    """uint8 [N,H]: code = #{k : word >= thr[k]}."""
    w = lev_words(seed, investor_ids, horizon).astype(np.uint64)
    code = np.zeros(w.shape, dtype=np.uint8)
    for t in thresholds(probs):
        code += (w >= np.uint64(t)).astype(np.uint8)
    return code


def box_muller_polar(a, b, sigma=1.0):
    """
    The engine's Box-Muller on two uint32 words (rlmd_b200/csrc/common.cuh):
    u1 = fl32(fl32(a) 2^-32 + 2^-33), theta = fl32(int32(b)) pi 2^-31,
    rho = sqrt(-2 ln2 sigma^2 log2(u1)); returns (rho cos theta, rho sin theta) in float64
    (the device uses lg2/sqrt/sin/cos.approx: agreement to ~1e-6).
    """
    a32 = np.asarray(a, dtype=np.uint32).astype(np.float32).astype(np.float64)     # RN conversion
    u1 = (a32 * 2.0 ** -32 + 2.0 ** -33).astype(np.float32).astype(np.float64)    # one rounding, like the FFMA
    b32 = np.asarray(b, dtype=np.uint32).view(np.int32).astype(np.float32)
    th = (b32 * np.float32(np.pi * 2.0 ** -31)).astype(np.float32).astype(np.float64)
    s32 = np.float32(sigma)
    scale2 = np.float64(np.float32(np.float32(np.float32(-1.3862943611198906) * s32) * s32))
    rho = np.sqrt(np.maximum(np.log2(u1) * scale2, 0.0))
    return rho * np.cos(th), rho * np.sin(th)


def gbm_returns(seed, investor_ids, horizon, log_mean, sigma) -> np.ndarray:
    """
    fp32 [N,H] (to ~1e-6: the device uses the approx lg2/sqrt/sin/cos units):
    block words (a,b,c,d) -> Box-Muller pairs (a,b) -> y0,y1 and (c,d) -> y2,y3 with
    sigma folded into the radius;  x = y + log_mean.
    """
    nblk = (horizon + 3) // 4
    w = lev_words(seed, investor_ids, nblk * 4).reshape(len(investor_ids), nblk, 4)
    y0, y1 = box_muller_polar(w[..., 0], w[..., 1], sigma)
    y2, y3 = box_muller_polar(w[..., 2], w[..., 3], sigma)
    x = np.stack([y0, y1, y2, y3], axis=-1) + np.float64(np.float32(log_mean))
    return x.reshape(len(investor_ids), nblk * 4)[:, :horizon].astype(np.float32)
