// SIMT validation kernels for the Vertex-RPE attention core (impl = 1 of vdetr_xattn_fwd/bwd) and the
// bias-only kernel.  These are deliberately simple fp32 kernels: they are the GPU-side cross-check for the
// tcgen05 product kernels at sizes where the CPU oracle is too slow, and the `return_attn_weights` debug path.
// They are NOT the product path (impl = 0 is).
//
// Math: /root/reference/models/vdetr_transformer.py:708-753.
#include "rpe_common.cuh"
#include "rpe_internal.h"

namespace {

constexpr int CH = 256;          // keys per chunk
constexpr int WARPS = 4;         // queries per CTA (one warp each)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct SimtArgs {
  VdetrXattnShape s;
  const float *q, *k, *v, *xyz, *ref, *ang;
  const float4* tables;
  float *out, *lse;
  // backward
  const float* dout;
  float *dq, *dk, *dv, *dtables;
};

// logits of one (b, q) row-group against key `key`: 4 heads
__device__ __forceinline__ float4 logits4(const SimtArgs& A, const float* qrow /*smem [H*hd]*/, const float* vert,
                                          float rc, float rs, int b, int key, const RpeParams& P) {
  const int hd = A.s.hd, kvh = A.s.kv_heads;
  float s[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float* kr = A.k + (((size_t)b * A.s.nK + key) * kvh + (kvh == 1 ? 0 : h)) * hd;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(qrow[h * hd + d], __ldg(kr + d), acc);
    s[h] = acc;
  }
  float4 r = make_float4(s[0], s[1], s[2], s[3]);
  if (A.s.has_bias) {
    const float* x = A.xyz + ((size_t)b * A.s.nK + key) * 3;
    float4 bias = rpe_bias_pair_exact(vert, x[0], x[1], x[2], rc, rs, A.tables, P);
    r.x += bias.x; r.y += bias.y; r.z += bias.z; r.w += bias.w;
  }
  return r;
}

__global__ void __launch_bounds__(WARPS * 32) simt_fwd_kernel(SimtArgs A) {
  __shared__ float sq[WARPS][256];
  __shared__ float sp[WARPS][4][CH];
  __shared__ float svert[WARPS][24];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS + warp;              // global (b, q)
  if (row >= A.s.B * A.s.nQ) return;
  const int b = row / A.s.nQ;
  const int hd = A.s.hd, nK = A.s.nK, kvh = A.s.kv_heads;
  RpeParams P{A.s.grid_n, A.s.log_scale, 1.0f / A.s.max_value, A.s.rotate};
  for (int i = lane; i < 4 * hd; i += 32) sq[warp][i] = A.q[(size_t)row * 4 * hd + i];
  if (A.s.has_bias && lane < 24) svert[warp][lane] = A.ref[(size_t)row * 24 + lane];
  float rc = 1.f, rs = 0.f;
  if (A.s.has_bias && A.s.rotate && A.ang) { float a = A.ang[row]; rc = cosf(a); rs = sinf(a); }
  __syncwarp();
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, l[4] = {0.f, 0.f, 0.f, 0.f};
  float acc[4][2] = {};                                   // lane owns output columns lane, lane+32 of each head
  for (int base = 0; base < nK; base += CH) {
    const int len = min(CH, nK - base);
    float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int t = lane; t < len; t += 32) {
      float4 s4 = logits4(A, sq[warp], svert[warp], rc, rs, b, base + t, P);
      sp[warp][0][t] = s4.x; sp[warp][1][t] = s4.y; sp[warp][2][t] = s4.z; sp[warp][3][t] = s4.w;
      cm[0] = fmaxf(cm[0], s4.x); cm[1] = fmaxf(cm[1], s4.y); cm[2] = fmaxf(cm[2], s4.z); cm[3] = fmaxf(cm[3], s4.w);
    }
    __syncwarp();
    float scale[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float mn = fmaxf(m[h], warp_max(cm[h]));
      scale[h] = (m[h] == -INFINITY) ? 0.f : expf(m[h] - mn);
      m[h] = mn;
      float part = 0.f;
      for (int t = lane; t < len; t += 32) {
        float p = expf(sp[warp][h][t] - mn);
        sp[warp][h][t] = p;
        part += p;
      }
      l[h] = l[h] * scale[h] + warp_sum(part);
      acc[h][0] *= scale[h]; acc[h][1] *= scale[h];
    }
    __syncwarp();
    for (int t = 0; t < len; ++t) {
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float* vr = A.v + (((size_t)b * nK + base + t) * kvh + (kvh == 1 ? 0 : h)) * hd;
        const float p = sp[warp][h][t];
        acc[h][0] = fmaf(p, __ldg(vr + lane), acc[h][0]);
        if (hd > 32) acc[h][1] = fmaf(p, __ldg(vr + lane + 32), acc[h][1]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float inv = 1.f / l[h];
    float* o = A.out + ((size_t)row * 4 + h) * hd;
    o[lane] = acc[h][0] * inv;
    if (hd > 32) o[lane + 32] = acc[h][1] * inv;
    if (lane == 0) A.lse[((size_t)b * 4 + h) * A.s.nQ + (row - b * A.s.nQ)] = m[h] + logf(l[h]);
  }
}

__global__ void __launch_bounds__(WARPS * 32) simt_bwd_kernel(SimtArgs A) {
  __shared__ float sq[WARPS][256];
  __shared__ float sdo[WARPS][256];
  __shared__ float sp[WARPS][4][CH];
  __shared__ float sds[WARPS][4][CH];
  __shared__ float svert[WARPS][24];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS + warp;
  if (row >= A.s.B * A.s.nQ) return;
  const int b = row / A.s.nQ, qi = row - b * A.s.nQ;
  const int hd = A.s.hd, nK = A.s.nK, kvh = A.s.kv_heads;
  RpeParams P{A.s.grid_n, A.s.log_scale, 1.0f / A.s.max_value, A.s.rotate};
  for (int i = lane; i < 4 * hd; i += 32) {
    sq[warp][i] = A.q[(size_t)row * 4 * hd + i];
    sdo[warp][i] = A.dout[(size_t)row * 4 * hd + i];
  }
  if (A.s.has_bias && lane < 24) svert[warp][lane] = A.ref[(size_t)row * 24 + lane];
  float rc = 1.f, rs = 0.f;
  if (A.s.has_bias && A.s.rotate && A.ang) { float a = A.ang[row]; rc = cosf(a); rs = sinf(a); }
  __syncwarp();
  float lse[4], delta[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    lse[h] = A.lse[((size_t)b * 4 + h) * A.s.nQ + qi];
    float part = 0.f;
    for (int d = lane; d < hd; d += 32) part += sdo[warp][h * hd + d] * A.out[((size_t)row * 4 + h) * hd + d];
    delta[h] = warp_sum(part);
  }
  float dq[4][2] = {};
  for (int base = 0; base < nK; base += CH) {
    const int len = min(CH, nK - base);
    for (int t = lane; t < len; t += 32) {
      const int key = base + t;
      float4 s4 = logits4(A, sq[warp], svert[warp], rc, rs, b, key, P);
      float p[4] = {expf(s4.x - lse[0]), expf(s4.y - lse[1]), expf(s4.z - lse[2]), expf(s4.w - lse[3])};
      float ds[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float* vr = A.v + (((size_t)b * nK + key) * kvh + (kvh == 1 ? 0 : h)) * hd;
        float dp = 0.f;
        for (int d = 0; d < hd; ++d) dp = fmaf(sdo[warp][h * hd + d], __ldg(vr + d), dp);
        ds[h] = p[h] * (dp - delta[h]);
        sp[warp][h][t] = p[h];
        sds[warp][h][t] = ds[h];
      }
      if (A.s.has_bias) {
        const float* x = A.xyz + ((size_t)b * nK + key) * 3;
        rpe_bias_pair_scatter(svert[warp], x[0], x[1], x[2], rc, rs, A.dtables, make_float4(ds[0], ds[1], ds[2], ds[3]), P);
      }
    }
    __syncwarp();
    for (int t = 0; t < len; ++t) {
      const int key = base + t;
      float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int kh = (kvh == 1 ? 0 : h);
        const float* kr = A.k + (((size_t)b * nK + key) * kvh + kh) * hd;
        const float ds = sds[warp][h][t], p = sp[warp][h][t];
        dq[h][0] = fmaf(ds, __ldg(kr + lane), dq[h][0]);
        if (hd > 32) dq[h][1] = fmaf(ds, __ldg(kr + lane + 32), dq[h][1]);
        if (kvh == 1) {
          dk0 = fmaf(ds, sq[warp][h * hd + lane], dk0);
          dv0 = fmaf(p, sdo[warp][h * hd + lane], dv0);
          if (hd > 32) { dk1 = fmaf(ds, sq[warp][h * hd + lane + 32], dk1); dv1 = fmaf(p, sdo[warp][h * hd + lane + 32], dv1); }
        } else {
          float* dkr = A.dk + (((size_t)b * nK + key) * kvh + h) * hd;
          float* dvr = A.dv + (((size_t)b * nK + key) * kvh + h) * hd;
          atomicAdd(dkr + lane, ds * sq[warp][h * hd + lane]);
          atomicAdd(dvr + lane, p * sdo[warp][h * hd + lane]);
          if (hd > 32) {
            atomicAdd(dkr + lane + 32, ds * sq[warp][h * hd + lane + 32]);
            atomicAdd(dvr + lane + 32, p * sdo[warp][h * hd + lane + 32]);
          }
        }
      }
      if (kvh == 1) {
        float* dkr = A.dk + ((size_t)b * nK + key) * hd;
        float* dvr = A.dv + ((size_t)b * nK + key) * hd;
        atomicAdd(dkr + lane, dk0); atomicAdd(dvr + lane, dv0);
        if (hd > 32) { atomicAdd(dkr + lane + 32, dk1); atomicAdd(dvr + lane + 32, dv1); }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float* o = A.dq + ((size_t)row * 4 + h) * hd;
    o[lane] = dq[h][0];
    if (hd > 32) o[lane + 32] = dq[h][1];
  }
}

__global__ void rpe_bias_kernel(VdetrXattnShape s, const float* __restrict__ xyz, const float* __restrict__ ref,
                                const float* __restrict__ ang, const float4* __restrict__ tables, float* __restrict__ rpe) {
  const size_t total = (size_t)s.B * s.nQ * s.nK;
  RpeParams P{s.grid_n, s.log_scale, 1.0f / s.max_value, s.rotate};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int key = (int)(i % s.nK);
    const size_t row = i / s.nK;                         // b*nQ + q
    const int b = (int)(row / s.nQ), q = (int)(row - (size_t)b * s.nQ);
    float vert[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) vert[j] = ref[row * 24 + j];
    float rc = 1.f, rs = 0.f;
    if (s.rotate && ang) { float a = ang[row]; rc = cosf(a); rs = sinf(a); }
    const float* x = xyz + ((size_t)b * s.nK + key) * 3;
    float4 v = rpe_bias_pair_exact(vert, x[0], x[1], x[2], rc, rs, tables, P);
    const size_t o = (((size_t)b * 4) * s.nQ + q) * s.nK + key;
    const size_t hs = (size_t)s.nQ * s.nK;
    rpe[o] = v.x; rpe[o + hs] = v.y; rpe[o + 2 * hs] = v.z; rpe[o + 3 * hs] = v.w;
  }
}

}  // namespace

int vdetr_check_shape(const VdetrXattnShape* s) {
  if (!s) return VDETR_ERR_BAD_ARG;
  if (s->B < 0 || s->nQ < 0 || s->nK < 0) return VDETR_ERR_BAD_ARG;
  if (s->H != 4 || s->hd != 64) return VDETR_ERR_UNSUPPORTED;
  if (s->kv_heads != 1 && s->kv_heads != s->H) return VDETR_ERR_UNSUPPORTED;
  if (s->has_bias && (s->grid_n < 2 || s->grid_n > 32 || !(s->max_value > 0.f))) return VDETR_ERR_BAD_ARG;
  return 0;
}

int simt_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                   const float* ref, const float* ang, const float* tables, float* out, float* lse, cudaStream_t st) {
  SimtArgs A = {};
  A.s = *s; A.q = q; A.k = k; A.v = v; A.xyz = xyz; A.ref = ref; A.ang = ang; A.tables = (const float4*)tables;
  A.out = out; A.lse = lse;
  const int rows = s->B * s->nQ;
  if (rows == 0) return 0;
  simt_fwd_kernel<<<(rows + WARPS - 1) / WARPS, WARPS * 32, 0, st>>>(A);
  VDETR_LAUNCH_CHECK();
  return 0;
}

int simt_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                   const float* ref, const float* ang, const float* tables, const float* out, const float* lse,
                   const float* dout, float* dq, float* dk, float* dv, float* dtables, cudaStream_t st) {
  SimtArgs A = {};
  A.s = *s; A.q = q; A.k = k; A.v = v; A.xyz = xyz; A.ref = ref; A.ang = ang; A.tables = (const float4*)tables;
  A.out = const_cast<float*>(out); A.lse = const_cast<float*>(lse); A.dout = dout;
  A.dq = dq; A.dk = dk; A.dv = dv; A.dtables = dtables;
  const size_t kvn = (size_t)s->B * s->nK * s->kv_heads * s->hd;
  VDETR_CUDA_TRY(cudaMemsetAsync(dk, 0, kvn * sizeof(float), st));
  VDETR_CUDA_TRY(cudaMemsetAsync(dv, 0, kvn * sizeof(float), st));
  if (s->has_bias)
    VDETR_CUDA_TRY(cudaMemsetAsync(dtables, 0, (size_t)8 * s->grid_n * s->grid_n * s->grid_n * 4 * sizeof(float), st));
  const int rows = s->B * s->nQ;
  if (rows == 0) return 0;
  simt_bwd_kernel<<<(rows + WARPS - 1) / WARPS, WARPS * 32, 0, st>>>(A);
  VDETR_LAUNCH_CHECK();
  return 0;
}

int rpe_bias_launch(const VdetrXattnShape* s, const float* xyz, const float* ref, const float* ang, const float* tables,
                    float* rpe, cudaStream_t st) {
  const size_t total = (size_t)s->B * s->nQ * s->nK;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  rpe_bias_kernel<<<blocks, 256, 0, st>>>(*s, xyz, ref, ang, (const float4*)tables, rpe);
  VDETR_LAUNCH_CHECK();
  return 0;
}
