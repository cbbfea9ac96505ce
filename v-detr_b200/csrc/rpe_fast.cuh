// Fast Vertex-RPE bias maths for the product kernels (forward, backward recompute, dTables scatter).
// Same formula as rpe_common.cuh / the reference (vdetr_transformer.py:708-731) with MUFU lg2 and a
// round-to-nearest floor; differences to the exact variant are ~1e-6 in the pixel coordinate.
#pragma once
#include "tc_common.cuh"

namespace rpe {

using tc::lg2_approx;
constexpr float MAGIC = 12582912.0f;              // 1.5 * 2^23
constexpr int MAGIC_BITS = 0x4B400000;

// ------------------------------------------------------------------------------------------------ bias maths
struct Axis {
  float w0, w1;     // weights of cell n0 and n0 + 1 (zero when the cell is padding)
  int n0;           // floor of the pixel coordinate, in [-2, n]
  int o0, o1;       // byte offsets of the two (clamped) cells along this axis
};
template <int STRIDE_SHIFT_UNUSED = 0>
__device__ __forceinline__ Axis rpe_axis_fast(float d, float ls, float c1, float c0, int n, int stride_bytes) {
  Axis a;
  float t = lg2_approx(fmaf(fabsf(d), ls, 1.0f)) * c1;
  float ts = copysignf(t, d);
  ts = fminf(fmaxf(ts, -c0 - 1.5f), (float)n - c0 + 0.5f);     // p in [-1.5, n+0.5]: outside both corners are padding
  float p = ts + c0;
  float r = (p - 0.5f) + MAGIC;                                 // round-to-nearest(p - 0.5) == floor(p) (ties are harmless)
  float fl = r - MAGIC;
  float f = p - fl;
  int n0 = __float_as_int(r) - MAGIC_BITS;
  a.n0 = n0;
  a.w0 = ((unsigned)n0 < (unsigned)n) ? 1.0f - f : 0.0f;
  a.w1 = ((unsigned)(n0 + 1) < (unsigned)n) ? f : 0.0f;
  a.o0 = min(max(n0, 0), n - 1) * stride_bytes;
  a.o1 = min(max(n0 + 1, 0), n - 1) * stride_bytes;
  return a;
}

__device__ __forceinline__ void corner8(float4& acc, const char* tab, const Axis& ax, const Axis& ay, const Axis& az) {
#pragma unroll
  for (int cz = 0; cz < 2; ++cz) {
    const int oz = cz ? az.o1 : az.o0;
    const float wz = cz ? az.w1 : az.w0;
#pragma unroll
    for (int cy = 0; cy < 2; ++cy) {
      const int ozy = oz + (cy ? ay.o1 : ay.o0);
      const float wzy = wz * (cy ? ay.w1 : ay.w0);
#pragma unroll
      for (int cx = 0; cx < 2; ++cx) {
        const float w = wzy * (cx ? ax.w1 : ax.w0);
        const float4 t = *reinterpret_cast<const float4*>(tab + ozy + (cx ? ax.o1 : ax.o0));
        acc.x = fmaf(w, t.x, acc.x); acc.y = fmaf(w, t.y, acc.y);
        acc.z = fmaf(w, t.z, acc.z); acc.w = fmaf(w, t.w, acc.w);
      }
    }
  }
}

// Bias of one (query,key) pair for the 4 heads.  geo: the query's record in shared memory (see pack kernel).
__device__ __forceinline__ float4 rpe_bias_pair(const float4* __restrict__ geo, float kx, float ky, float kz,
                                                const char* __restrict__ tab, int n, float ls, float c1, float c0) {
  const int sx = 16, sy = 16 * n, sz = 16 * n * n, st = 16 * n * n * n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 hi = geo[0];
  if (__float_as_int(hi.w) != 0) {
    // axis-aligned box: 2 distinct coordinates per axis -> 6 transforms instead of 24
    const float4 lo = geo[1];
    const Axis xp = rpe_axis_fast(hi.x - kx, ls, c1, c0, n, sx), xm = rpe_axis_fast(lo.x - kx, ls, c1, c0, n, sx);
    const Axis yp = rpe_axis_fast(hi.y - ky, ls, c1, c0, n, sy), ym = rpe_axis_fast(lo.y - ky, ls, c1, c0, n, sy);
    const Axis zp = rpe_axis_fast(hi.z - kz, ls, c1, c0, n, sz), zm = rpe_axis_fast(lo.z - kz, ls, c1, c0, n, sz);
    // vertex sign table: 0:(+,+,-) 1:(+,-,-) 2:(-,-,-) 3:(-,+,-) 4:(+,+,+) 5:(+,-,+) 6:(-,-,+) 7:(-,+,+)
    corner8(acc, tab + 0 * st, xp, yp, zm);
    corner8(acc, tab + 1 * st, xp, ym, zm);
    corner8(acc, tab + 2 * st, xm, ym, zm);
    corner8(acc, tab + 3 * st, xm, yp, zm);
    corner8(acc, tab + 4 * st, xp, yp, zp);
    corner8(acc, tab + 5 * st, xp, ym, zp);
    corner8(acc, tab + 6 * st, xm, ym, zp);
    corner8(acc, tab + 7 * st, xm, yp, zp);
  } else {
    const float4 rot = geo[8];
    const float* v = reinterpret_cast<const float*>(geo + 2);
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      float dx = v[i * 3 + 0] - kx, dy = v[i * 3 + 1] - ky, dz = v[i * 3 + 2] - kz;
      const float tx = rot.x * dx - rot.y * dy, ty = rot.y * dx + rot.x * dy;     // identity when not rotated
      const Axis ax = rpe_axis_fast(tx, ls, c1, c0, n, sx), ay = rpe_axis_fast(ty, ls, c1, c0, n, sy),
                 az = rpe_axis_fast(dz, ls, c1, c0, n, sz);
      corner8(acc, tab + i * st, ax, ay, az);
    }
  }
  return acc;
}


}  // namespace rpe
