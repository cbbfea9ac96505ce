// AdamW on flat parameter / gradient / moment buffers: ONE launch per step instead of the ~22 multi-tensor launches (0.95 ms)
// torch.optim.AdamW(fused=True) needs for the decoder's ~1000 small tensors (the reference: optimizer.py:25, stepped in
// engine.py:105-108 after clip_grad_norm_).  Same update as torch.optim.AdamW:
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// lr, the step count t and an optional gradient scale (1 / world size, gradient clipping) are read from device memory, so a
// captured CUDA graph follows a learning-rate schedule without being re-captured.  Elements [0, n_decay) are weight-decayed,
// the rest (biases / norm parameters under --filter_biases_wd, optimizer.py:11-22) are not.
#include "common.cuh"

namespace {

struct AdamParams {
  float4* p; const float4* g; float4* m; float4* v;
  long long n4, n, n_decay;
  const float* lr; const float* step; const float* gscale;
  float b1, b2, eps, wd, gscale_host;
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float decay, float b1, float b2, float eps, float step_size,
                                      float inv_sqrt_bc2) {
  p *= decay;
  m = m + (1.f - b1) * (g - m);                 // lerp
  v = v * b2 + (1.f - b2) * g * g;
  p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
}

__global__ void __launch_bounds__(256) adamw_flat_kernel(const AdamParams A) {
  const float lr = __ldg(A.lr), t = __ldg(A.step);
  const float gs = A.gscale_host * (A.gscale ? __ldg(A.gscale) : 1.f);
  const float bc1 = 1.f - powf(A.b1, t), bc2 = 1.f - powf(A.b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay_on = 1.f - lr * A.wd;
  float* ps = reinterpret_cast<float*>(A.p);
  float* ms = reinterpret_cast<float*>(A.m);
  float* vs = reinterpret_cast<float*>(A.v);
  const float* gsrc = reinterpret_cast<const float*>(A.g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = A.p[i], m = A.m[i], v = A.v[i];
    float4 g = __ldg(A.g + i);
    g.x *= gs; g.y *= gs; g.z *= gs; g.w *= gs;
    const long long e = i * 4;
    adam1(p.x, g.x, m.x, v.x, e + 0 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    adam1(p.y, g.y, m.y, v.y, e + 1 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    adam1(p.z, g.z, m.z, v.z, e + 2 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    adam1(p.w, g.w, m.w, v.w, e + 3 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    A.p[i] = p; A.m[i] = m; A.v[i] = v;
  }
  // tail (n not a multiple of 4)
  const long long tail = A.n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && tail < A.n) {
    float p = ps[tail], m = ms[tail], v = vs[tail];
    adam1(p, gsrc[tail] * gs, m, v, tail < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    ps[tail] = p; ms[tail] = m; vs[tail] = v;
  }
}

}  // namespace

extern "C" int vdetr_adamw_flat(float* p, const float* g, float* m, float* v, long long n, long long n_decay, const float* lr,
                                const float* step, const float* grad_scale, float grad_scale_host, float beta1, float beta2,
                                float eps, float weight_decay, void* stream) {
  if (n < 0 || n_decay < 0 || n_decay > n) return VDETR_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!p || !g || !m || !v || !lr || !step) return VDETR_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15)
    return VDETR_ERR_BAD_ARG;
  AdamParams A;
  A.p = reinterpret_cast<float4*>(p); A.g = reinterpret_cast<const float4*>(g);
  A.m = reinterpret_cast<float4*>(m); A.v = reinterpret_cast<float4*>(v);
  A.n = n; A.n4 = n / 4; A.n_decay = n_decay;
  A.lr = lr; A.step = step; A.gscale = grad_scale; A.gscale_host = grad_scale_host;
  A.b1 = beta1; A.b2 = beta2; A.eps = eps; A.wd = weight_decay;
  long long blocks = (A.n4 + 255) / 256;
  const long long cap = (long long)vdetr_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adamw_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(A);
  VDETR_LAUNCH_CHECK();
  return 0;
}
