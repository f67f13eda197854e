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


def gbm_returns(seed, investor_ids, horizon, log_mean, sigma) -> np.ndarray:
    """
    fp32 [N,H] (to ~1e-6: the device uses fast log/sin/cos intrinsics):
    block words (a,b,c,d) -> Box-Muller pairs (a,b) -> z0,z1 and (c,d) -> z2,z3,
    u1 = ((a >> 8) + 0.5) 2^-24, u2 = (b >> 8) 2^-24, r = sqrt(-2 ln u1),
    z = r cos(2 pi u2), r sin(2 pi u2);  x = sigma z + log_mean.
    """
    nblk = (horizon + 3) // 4
    w = lev_words(seed, investor_ids, nblk * 4).reshape(len(investor_ids), nblk, 4)
    u1a = ((w[..., 0] >> 8).astype(np.float64) + 0.5) * 2.0 ** -24
    u2a = (w[..., 1] >> 8).astype(np.float64) * 2.0 ** -24
    u1b = ((w[..., 2] >> 8).astype(np.float64) + 0.5) * 2.0 ** -24
    u2b = (w[..., 3] >> 8).astype(np.float64) * 2.0 ** -24
    ra, rb = np.sqrt(-2 * np.log(u1a)), np.sqrt(-2 * np.log(u1b))
    z = np.stack([ra * np.cos(2 * np.pi * u2a), ra * np.sin(2 * np.pi * u2a),
                  rb * np.cos(2 * np.pi * u2b), rb * np.sin(2 * np.pi * u2b)], axis=-1)
    x = np.float64(np.float32(sigma)) * z + np.float64(np.float32(log_mean))
    return x.reshape(len(investor_ids), nblk * 4)[:, :horizon].astype(np.float32)
