"""Helpers shared by the pointnet2 tests: ctypes access to the C oracle and synthetic point clouds."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def oracle_lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_build", "libpn2_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libpn2_oracle.so"])
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def ref_fps(xyz, m):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    b, n, _ = xyz.shape
    out = np.zeros((b, m), dtype=np.int32)
    oracle_lib().pn2_ref_fps(b, n, m, _p(xyz, ctypes.c_float), _p(out, ctypes.c_int32))
    return out


def ref_ball_query(new_xyz, xyz, radius, nsample):
    new_xyz = np.ascontiguousarray(new_xyz, dtype=np.float32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = np.zeros((b, m, nsample), dtype=np.int32)
    oracle_lib().pn2_ref_ball_query(b, n, m, ctypes.c_float(radius), nsample, _p(new_xyz, ctypes.c_float),
                                    _p(xyz, ctypes.c_float), _p(out, ctypes.c_int32))
    return out


def lattice_cloud(seed, b, n, zero_frac=0.0):
    """ScanNet-like: room 8x8x3 m on the 0.04 m voxel lattice (exact distance ties are common)."""
    rs = np.random.RandomState(seed)
    pts = (np.round(rs.rand(b, n, 3) * np.array([8.0, 8.0, 3.0]) / 0.04) * 0.04).astype(np.float32)
    if zero_frac > 0:
        m = rs.rand(b, n) < zero_frac
        pts[m] = (rs.rand(int(m.sum()), 3) * 0.015).astype(np.float32)     # |p|^2 <= 1e-3 -> skipped by FPS
    return pts


def fps_slot_order_numpy(xyz, m):
    """FPS restated through the 'slot' total order used by the CUDA kernel (v-detr_b200/csrc/pointnet2.cu):
    winner = argmax temp, ties -> smallest slot(k) = bitrev(k mod S)*ceil(n/S) + k div S."""
    xyz = np.asarray(xyz, dtype=np.float32)
    b, n, _ = xyz.shape
    pow2 = int(np.log(float(n)) / np.log(2.0))
    S = max(min(1 << pow2, 512), 1)
    L = S.bit_length() - 1
    cnt = (n + S - 1) // S
    k = np.arange(n)
    low = k & (S - 1)
    rev = np.zeros_like(low)
    for i in range(L):
        rev |= ((low >> i) & 1) << (L - 1 - i)
    slot = rev * cnt + (k >> L)
    out = np.zeros((b, m), dtype=np.int32)
    for bi in range(b):
        p = xyz[bi]
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        f = np.float32
        mag = (z * z + (y * y + (x * x).astype(f)).astype(f)).astype(f)      # not fma-exact; only used for the skip rule
        mag64 = x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2 + z.astype(np.float64) ** 2
        # exact fma emulation in float64: fma(c,c,fma(b,b,a*a)) -- a*a rounded to f32, then each fma rounded once
        def sumsq3(a, bb, c):      # fma(z,z, fma(x,x, y*y)): the contraction nvcc emits for the reference
            t = (bb.astype(np.float64) * bb.astype(np.float64)).astype(f)
            t = (a.astype(np.float64) * a.astype(np.float64) + t.astype(np.float64)).astype(f)
            return (c.astype(np.float64) * c.astype(np.float64) + t.astype(np.float64)).astype(f)
        mag = sumsq3(x, y, z)
        _ = mag64
        valid = ~(mag.astype(np.float64) <= 1e-3)
        temp = np.full(n, 1e10, dtype=f)
        old = 0
        for j in range(1, m):
            d = sumsq3((x - x[old]).astype(f), (y - y[old]).astype(f), (z - z[old]).astype(f))
            temp = np.where(valid, np.minimum(d, temp), temp)
            if not valid.any():
                old = 0
            else:
                tv = np.where(valid, temp, -1.0)
                best = tv.max()
                cand = np.nonzero(tv == best)[0]
                old = int(cand[np.argmin(slot[cand])])
            out[bi, j] = old
    return out
