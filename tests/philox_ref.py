"""numpy restatement of csrc/philox.cuh (Philox4x32-10 keep mask of the fused attention dropout).  Test infrastructure."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & MASK32
        k1 = (k1 + np.uint64(W1)) & MASK32
    return c0, c1, c2, c3


def thresh_of(p):
    return int(min(max(np.float32(p) * np.float32(65536.0) + np.float32(0.5), 0.0), 65535.0))


def keep_mask(seed, B, H, nQ, nK, p, kvh=1):
    """bool [B,H,nQ,nK]: element kept by the fused kernels for this seed (packed row index as in rpe_internal.h)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    nQp = (nQ + 31) // 32 * 32 if kvh == 1 else (nQ + 127) // 128 * 128
    b, h, q, key = np.meshgrid(np.arange(B), np.arange(H), np.arange(nQ), np.arange(nK), indexing="ij")
    row = (b * nQp + q) * 4 + h if kvh == 1 else (b * 4 + h) * nQp + q
    r = philox4x32_10(key >> 3, row, np.zeros_like(row), np.zeros_like(row), k0, k1)
    w = np.choose((key & 7) >> 1, r)
    val = np.where(key & 1, w >> np.uint64(16), w & np.uint64(0xFFFF))
    return val >= np.uint64(thresh_of(p))
