// Fast Vertex-RPE bias maths for the product kernels (forward, backward recompute, dTables scatter).
// Same formula as rpe_common.cuh / the reference (vdetr_transformer.py:708-731) with MUFU lg2 and a
// round-to-nearest floor; differences to the exact variant are ~1e-6 in the pixel coordinate.
#pragma once
#include <cuda_fp16.h>
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

// ---- table layout in shared memory -----------------------------------------------------------------------
// The fused kernels are bound by shared-memory wavefronts (an LDS.128 always costs 4, broadcast or not), so the
// tables are stored as FP16 *x-pairs*: entry (i, z, y, p), p in [0, n-2], holds the 4 heads of cells x = p and
// x = p + 1 (8 halves = 16 B).  One LDS.128 then serves both x-corners of a (z, y) corner pair: 32 loads per
// (query, key) pair instead of 64.  Cells outside [0, n-1] are padding (weight 0), so x0 = -1 maps to pair 0 with
// the weight on its low element and x0 = n-1 to pair n-2 with the weight on its high element.
__host__ __device__ inline int pair_table_bytes(int n) { return 8 * n * n * (n - 1) * 16; }

struct XPair {
  float wl, wh;     // weights of the low / high cell of the pair
  int off;          // byte offset of the pair inside a (z, y) row
};
__device__ __forceinline__ XPair make_xpair(const Axis& a, int n) {
  XPair x;
  const bool lt = a.n0 < 0, gt = a.n0 > n - 2;
  x.wl = lt ? a.w1 : (gt ? 0.0f : a.w0);
  x.wh = lt ? 0.0f : (gt ? a.w0 : a.w1);
  x.off = min(max(a.n0, 0), n - 2) * 16;
  return x;
}

// d = a * b + c with FP16 multiplicands and an FP32 addend / result in ONE instruction (FHFMA, new on sm_100): the
// table halves can be consumed as they come out of shared memory, without the 8 HADD2.F32 conversions per LDS.128 that
// make up a quarter of the fused kernels' instructions -- but the interpolation weight has to be rounded to FP16 too.
// Measured on B200 (B = 8 x 1024 x 4096): fused forward 1.25 -> 0.99 ms, bias rounding noise x1.4 (3.2e-4 -> 4.6e-4
// rms at |T| <= 4), which the deliberately chaotic train golden case amplifies to 6.5 % median gradient difference
// (3.8 % with FP32 weights).  Parity first: the default keeps FP32 weights; -DRPE_FHFMA=1 is the opt-in.
#ifndef RPE_FHFMA
#define RPE_FHFMA 0
#endif
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void split_b32(unsigned x, unsigned short& lo, unsigned short& hi) {
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(x));
}

// One vertex: 4 (z, y) corners x one x-pair entry.  tabx = vertex table + byte offset of the x pair; row[j] / wzy[j] =
// byte offset and weight product of (z, y) corner j (shared by the two vertices that differ only in the x sign).
__device__ __forceinline__ void vertex_pairs(float4& acc, const char* tabx, const XPair& x, const int (&row)[4], const float (&wzy)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 raw = *reinterpret_cast<const uint4*>(tabx + row[j]);
    const float a = wzy[j] * x.wl, b = wzy[j] * x.wh;
#if RPE_FHFMA
    const __half2 ab = __floats2half2_rn(a, b);
    unsigned short a16, b16, l0, l1, l2, l3, h0, h1, h2, h3;
    split_b32(*reinterpret_cast<const unsigned*>(&ab), a16, b16);
    split_b32(raw.x, l0, l1); split_b32(raw.y, l2, l3); split_b32(raw.z, h0, h1); split_b32(raw.w, h2, h3);
    acc.x = fhfma(l0, a16, fhfma(h0, b16, acc.x)); acc.y = fhfma(l1, a16, fhfma(h1, b16, acc.y));
    acc.z = fhfma(l2, a16, fhfma(h2, b16, acc.z)); acc.w = fhfma(l3, a16, fhfma(h3, b16, acc.w));
#else
    const float2 l01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 l23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.z));
    const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.w));
    acc.x = fmaf(a, l01.x, fmaf(b, h01.x, acc.x)); acc.y = fmaf(a, l01.y, fmaf(b, h01.y, acc.y));
    acc.z = fmaf(a, l23.x, fmaf(b, h23.x, acc.z)); acc.w = fmaf(a, l23.y, fmaf(b, h23.y, acc.w));
#endif
  }
}
__device__ __forceinline__ void zy_corners(const Axis& ay, const Axis& az, int (&row)[4], float (&wzy)[4]) {
#pragma unroll
  for (int cz = 0; cz < 2; ++cz)
#pragma unroll
    for (int cy = 0; cy < 2; ++cy) {
      row[cz * 2 + cy] = (cz ? az.o1 : az.o0) + (cy ? ay.o1 : ay.o0);
      wzy[cz * 2 + cy] = (cz ? az.w1 : az.w0) * (cy ? ay.w1 : ay.w0);
    }
}

// fp32 tables [8][n][n][n] float4 (global) -> fp16 x-pair tables (shared); cooperative, `nthreads` participants
__device__ __forceinline__ void load_pair_tables(uint4* dst, const float4* __restrict__ src, int n, int tid, int nthreads) {
  const int npair = n - 1, total = 8 * n * n * npair;
  for (int i = tid; i < total; i += nthreads) {
    const int p = i % npair, row = i / npair;
    const float4 lo = __ldg(src + row * n + p), hi = __ldg(src + row * n + p + 1);
    uint4 o;
    o.x = tc::pack_f16x2(lo.x, lo.y); o.y = tc::pack_f16x2(lo.z, lo.w);
    o.z = tc::pack_f16x2(hi.x, hi.y); o.w = tc::pack_f16x2(hi.z, hi.w);
    dst[i] = o;
  }
}

// Bias of one (query,key) pair for the 4 heads.  geo: the query's record in shared memory (see pack kernel);
// tab: fp16 x-pair tables in shared memory.
__device__ __forceinline__ float4 rpe_bias_pair(const float4* __restrict__ geo, float kx, float ky, float kz,
                                                const char* __restrict__ tab, int n, float ls, float c1, float c0) {
  const int sx = 16, sy = 16 * (n - 1), sz = 16 * n * (n - 1), st = 16 * n * n * (n - 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 hi = geo[0];
  if (__float_as_int(hi.w) != 0) {
    // axis-aligned box: 2 distinct coordinates per axis -> 6 transforms instead of 24
    const float4 lo = geo[1];
    const XPair xp = make_xpair(rpe_axis_fast(hi.x - kx, ls, c1, c0, n, sx), n);
    const XPair xm = make_xpair(rpe_axis_fast(lo.x - kx, ls, c1, c0, n, sx), n);
    const Axis yp = rpe_axis_fast(hi.y - ky, ls, c1, c0, n, sy), ym = rpe_axis_fast(lo.y - ky, ls, c1, c0, n, sy);
    const Axis zp = rpe_axis_fast(hi.z - kz, ls, c1, c0, n, sz), zm = rpe_axis_fast(lo.z - kz, ls, c1, c0, n, sz);
    const char* tp = tab + xp.off;
    const char* tm = tab + xm.off;
    int row[4];
    float wzy[4];
    // vertex sign table: 0:(+,+,-) 1:(+,-,-) 2:(-,-,-) 3:(-,+,-) 4:(+,+,+) 5:(+,-,+) 6:(-,-,+) 7:(-,+,+)
    // the two vertices of a line differ only in the x sign and share the (z, y) corner offsets and weights
    zy_corners(yp, zm, row, wzy); vertex_pairs(acc, tp + 0 * st, xp, row, wzy); vertex_pairs(acc, tm + 3 * st, xm, row, wzy);
    zy_corners(ym, zm, row, wzy); vertex_pairs(acc, tp + 1 * st, xp, row, wzy); vertex_pairs(acc, tm + 2 * st, xm, row, wzy);
    zy_corners(yp, zp, row, wzy); vertex_pairs(acc, tp + 4 * st, xp, row, wzy); vertex_pairs(acc, tm + 7 * st, xm, row, wzy);
    zy_corners(ym, zp, row, wzy); vertex_pairs(acc, tp + 5 * st, xp, row, wzy); vertex_pairs(acc, tm + 6 * st, xm, row, wzy);
  } else {
    const float4 rot = geo[8];
    const float* v = reinterpret_cast<const float*>(geo + 2);
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      float dx = v[i * 3 + 0] - kx, dy = v[i * 3 + 1] - ky, dz = v[i * 3 + 2] - kz;
      const float tx = rot.x * dx - rot.y * dy, ty = rot.y * dx + rot.x * dy;     // identity when not rotated
      const XPair ax = make_xpair(rpe_axis_fast(tx, ls, c1, c0, n, sx), n);
      const Axis ay = rpe_axis_fast(ty, ls, c1, c0, n, sy), az = rpe_axis_fast(dz, ls, c1, c0, n, sz);
      int row[4];
      float wzy[4];
      zy_corners(ay, az, row, wzy);
      vertex_pairs(acc, tab + i * st + ax.off, ax, row, wzy);
    }
  }
  return acc;
}

}  // namespace rpe
