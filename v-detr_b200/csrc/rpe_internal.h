// Internal (non-ABI) entry points shared between the translation units of libvdetr_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "../../include/vdetr_b200.h"

int vdetr_check_shape(const VdetrXattnShape* s);

// impl = 1 (validation kernels, rpe_simt.cu)
int simt_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                   const float* ref, const float* ang, const float* tables, float* out, float* lse, cudaStream_t st);
int simt_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                   const float* ref, const float* ang, const float* tables, const float* out, const float* lse,
                   const float* dout, float* dq, float* dk, float* dv, float* dtables, cudaStream_t st);
int rpe_bias_launch(const VdetrXattnShape* s, const float* xyz, const float* ref, const float* ang, const float* tables,
                    float* rpe, cudaStream_t st);

// impl = 0 (product kernels: tcgen05 + TMA, rpe_xattn_fwd.cu / rpe_xattn_bwd.cu)
size_t tc_xattn_fwd_workspace(const VdetrXattnShape* s);
// drop_p > 0 with a device seed: dropout on the attention probabilities (philox.cuh), identical masks in fwd / bwd
int tc_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                 const float* ref, const float* ang, const float* tables, float* out, float* lse, float* bias_save,
                 float drop_p, const unsigned long long* drop_seed, void* ws, size_t ws_bytes, cudaStream_t st);
// bytes of the optional per-pair bias buffer the forward can leave for the backward ([B][nQp][nKp] float4; 0 = n/a)
size_t tc_xattn_bias_save_bytes(const VdetrXattnShape* s);
// bias_is_saved = 0: room for re-running the forward kernel into a transient bias buffer is included
size_t tc_xattn_bwd_workspace(const VdetrXattnShape* s, int bias_is_saved);
int tc_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                 const float* ref, const float* ang, const float* tables, const float* out, const float* lse,
                 const float* dout, const float* bias_saved, float drop_p, const unsigned long long* drop_seed, float* dq,
                 float* dk, float* dv, float* dtables, void* ws, size_t ws_bytes, cudaStream_t st);

// Operand packing shared by the product forward / backward kernels (rpe_xattn_fwd.cu):
//   qp   fp16 [rows][64]   MQA rows = (b*nQp + q)*4 + h        MHA rows = (b*4 + h)*nQp + q     (zero padded)
//   qpl  fp16, same layout: the rounding residual q - fp16(q) (hi/lo split: S = Qh Kh^T + Ql Kh^T + Qh Kl^T is
//        accurate to ~2^-22 relative although every MMA operand is fp16)
//   dop  fp16, same layout as qp (backward only)
//   kp / kpl fp16 [B][kvh][nKp][64] (hi / lo)     vp fp16 same layout (backward only)
//   vtp  fp16 [B][kvh][64][nKp]   (V transposed: keys contiguous)
//   xyz4 f32  [B][nKp] float4
//   geo  f32  [B][nQp][9] float4: (x+,y+,z+,fast flag) (x-,y-,z-,0) 24 vertex floats (cos,sin,0,0)
// Precision plan: every tensor-core operand is FP16 (11-bit significand: 8x tighter than BF16 at the same
// cost).  q/k/v/p are O(1), so range is no concern for them; gradients (dO, and dS which is proportional to it)
// are multiplied by a power of two chosen per call from max|dO| so that they sit in [2^-20, 16] of fp16's range
// ("scaled fp16"), and the results are divided by the same power of two (exact).
struct VdetrPack {
  int B, nQ, nK, nQp, nKp, kvh, has_bias;
  const float *q, *k, *v, *xyz, *ref, *ang, *dout;
  const unsigned* dout_absmax_bits;      // device: bits of max|dout| (backward only)
  __half *qp, *kp, *vtp;                 // forward operands (and the S recompute of the backward)
  __half *qpl, *kpl;                     // rounding residuals of qp / kp (hi/lo split of the S = Q K^T operands)
  __half *vp, *dop;                      // backward: row-major V, scaled dO (null in the forward)
  float4* xyz4;
  float4* geo;
};
__global__ void vdetr_pack_kernel(VdetrPack K);
__global__ void vdetr_absmax_kernel(const float* x, size_t n, unsigned* out_bits);
// power-of-two gradient scale derived from max|dout| (same formula on every device thread that needs it)
__device__ __forceinline__ float vdetr_grad_scale(unsigned absmax_bits) {
  const float m = __uint_as_float(absmax_bits);
  if (!(m > 0.f) || !(m < 3.0e38f)) return 1.f;
  float e = floorf(log2f(16.f / m));
  e = fminf(fmaxf(e, -100.f), 100.f);
  return exp2f(e);
}

// scale of the fp16 dS rows the dTables kernels read: the backward's gradient scale, or (dense C-ABI helper) its own rule
__device__ __forceinline__ float vdetr_dt_scale(unsigned bits, int dense) {
  if (!dense) return vdetr_grad_scale(bits);
  const float m = __uint_as_float(bits);
  if (!(m > 0.f) || !(m < 3.0e38f)) return 1.f;
  float e = floorf(log2f(16384.f / m));
  e = fminf(fmaxf(e, -100.f), 100.f);
  return exp2f(e);
}

// dTables of axis-aligned boxes as a dense tcgen05 contraction (rpe_dtables_umma.cu)
size_t rpe_dt6_priv_bytes();
int rpe_dt6_launch(const VdetrXattnShape* s, int nQp, int nKp, const float4* xyz4, const float4* geo, const __half* dsb,
                   const unsigned* absmax_bits, int dense_scale, float* dtables, int accumulate, float* priv, cudaStream_t st);

// dTables (rpe_dtables.cu)
size_t rpe_dtables_scratch_bytes(const VdetrXattnShape* s);
int rpe_dtables_launch(const VdetrXattnShape* s, int nQp, int nKp, const float4* xyz4, const float4* geo, const __half* dsb,
                       const unsigned* absmax_bits, int dense_scale, float* dtables, void* scratch, size_t scratch_bytes,
                       cudaStream_t st);
size_t rpe_dtables_workspace(const VdetrXattnShape* s);
int rpe_dtables_dense(const VdetrXattnShape* s, const float* xyz, const float* ref, const float* ang, const float* ds4,
                      float* dtables, void* ws, size_t ws_bytes, cudaStream_t st);
