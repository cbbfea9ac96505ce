// Vertex-RPE bias maths shared by the kernels.  Follows /root/reference/models/vdetr_transformer.py:708-731
// (SURVEY.md Appendix A steps 3-7): delta -> optional per-query rotation -> signed log2 -> align_corners=False
// pixel coordinate -> 8-corner trilinear gather with zero padding from tables[i][z][y][x][h].
#pragma once
#include "common.cuh"

struct RpeParams {
  int n;             // table points per axis
  float log_scale;   // 512
  float inv_max;     // 1 / max_value
  int rotate;
};

// g(d) * n/2 + (n-1)/2 written in the reference's operation order (accurate variant, used by the
// validation kernels): sign(d)*log2(|d|*ls+1)/3/max ; p = ((g+1)*n-1)/2.
__device__ __forceinline__ float rpe_pixel_exact(float d, const RpeParams& P) {
  float a = log2f(fmaf(fabsf(d), P.log_scale, 1.0f));
  float g = copysignf(a, d);
  if (d == 0.0f) g = 0.0f;                 // torch.sign(0) = 0
  g = g / 3.0f * P.inv_max;
  return ((g + 1.0f) * (float)P.n - 1.0f) * 0.5f;
}

struct RpeAxis {   // one axis of one vertex: base index and the two (zero-padded) weights
  int i0;
  float w0, w1;
};
__device__ __forceinline__ RpeAxis rpe_axis(float p, int n) {
  RpeAxis a;
  // clamp so that the float->int conversion is defined; outside [-1, n) both corners are out of range
  float pc = fminf(fmaxf(p, -2.0f), (float)n + 1.0f);
  float fl = floorf(pc);
  float f = pc - fl;
  a.i0 = (int)fl;
  a.w0 = (a.i0 >= 0 && a.i0 < n) ? 1.0f - f : 0.0f;
  a.w1 = (a.i0 + 1 >= 0 && a.i0 + 1 < n) ? f : 0.0f;
  if (!(p == p)) { a.w0 = a.w1 = 0.f; a.i0 = 0; }   // NaN coordinates contribute nothing
  return a;
}

// Bias for one (query, key) pair, all H=4 heads.  vert: 8 vertices x (x,y,z); tables: [8][n][n][n] float4.
__device__ __forceinline__ float4 rpe_bias_pair_exact(const float* __restrict__ vert, float kx, float ky, float kz,
                                                      float rc, float rs, const float4* __restrict__ tables,
                                                      const RpeParams& P) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n = P.n;
  const int n3 = n * n * n;
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    float dx = vert[i * 3 + 0] - kx, dy = vert[i * 3 + 1] - ky, dz = vert[i * 3 + 2] - kz;
    if (P.rotate) {
      float tx = rc * dx - rs * dy, ty = rs * dx + rc * dy;
      dx = tx; dy = ty;
    }
    RpeAxis ax = rpe_axis(rpe_pixel_exact(dx, P), n);
    RpeAxis ay = rpe_axis(rpe_pixel_exact(dy, P), n);
    RpeAxis az = rpe_axis(rpe_pixel_exact(dz, P), n);
    const float4* T = tables + (size_t)i * n3;
#pragma unroll
    for (int cz = 0; cz < 2; ++cz) {
      const float wz = cz ? az.w1 : az.w0;
      const int iz = min(max(az.i0 + cz, 0), n - 1);
#pragma unroll
      for (int cy = 0; cy < 2; ++cy) {
        const float wzy = wz * (cy ? ay.w1 : ay.w0);
        const int iy = min(max(ay.i0 + cy, 0), n - 1);
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
          const float w = wzy * (cx ? ax.w1 : ax.w0);
          const int ix = min(max(ax.i0 + cx, 0), n - 1);
          const float4 t = __ldg(T + (iz * n + iy) * n + ix);
          acc.x = fmaf(w, t.x, acc.x); acc.y = fmaf(w, t.y, acc.y);
          acc.z = fmaf(w, t.z, acc.z); acc.w = fmaf(w, t.w, acc.w);
        }
      }
    }
  }
  return acc;
}

// Adjoint: scatter ds (4 heads) of one pair into dtables with global atomics (validation kernels only).
__device__ __forceinline__ void rpe_bias_pair_scatter(const float* __restrict__ vert, float kx, float ky, float kz,
                                                      float rc, float rs, float* __restrict__ dtables, float4 ds,
                                                      const RpeParams& P) {
  const int n = P.n;
  const int n3 = n * n * n;
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    float dx = vert[i * 3 + 0] - kx, dy = vert[i * 3 + 1] - ky, dz = vert[i * 3 + 2] - kz;
    if (P.rotate) {
      float tx = rc * dx - rs * dy, ty = rs * dx + rc * dy;
      dx = tx; dy = ty;
    }
    RpeAxis ax = rpe_axis(rpe_pixel_exact(dx, P), n);
    RpeAxis ay = rpe_axis(rpe_pixel_exact(dy, P), n);
    RpeAxis az = rpe_axis(rpe_pixel_exact(dz, P), n);
    float* T = dtables + (size_t)i * n3 * 4;
    for (int c = 0; c < 8; ++c) {
      const int cz = c >> 2, cy = (c >> 1) & 1, cx = c & 1;
      const float w = (cz ? az.w1 : az.w0) * (cy ? ay.w1 : ay.w0) * (cx ? ax.w1 : ax.w0);
      if (w == 0.f) continue;
      const int iz = min(max(az.i0 + cz, 0), n - 1), iy = min(max(ay.i0 + cy, 0), n - 1),
                ix = min(max(ax.i0 + cx, 0), n - 1);
      float* cell = T + (size_t)((iz * n + iy) * n + ix) * 4;
      atomicAdd(cell + 0, w * ds.x); atomicAdd(cell + 1, w * ds.y);
      atomicAdd(cell + 2, w * ds.z); atomicAdd(cell + 3, w * ds.w);
    }
  }
}
