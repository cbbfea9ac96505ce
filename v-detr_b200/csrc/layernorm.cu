// LayerNorm over the last dimension, forward and backward, for the decoder's [tokens, 256] activations
// (nn.LayerNorm in GlobalDecoderLayer / FFNLayer / TransformerDecoder.norm: models/vdetr_transformer.py:463-466,
// 586-606, 129).  One warp per row, the whole row in registers (cols = 128 * VEC, VEC float4 per lane).
//   forward : y = (x - mean) * rstd * gamma + beta, mean / rstd saved
//   backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; dgamma / dbeta are accumulated per
//             lane over the rows a warp walks, reduced across the CTA in shared memory, stored as per-CTA partial sums
//             and added up in block order by the last CTA (det_reduce.cuh; the 34 LayerNorm backward calls of a decoder
//             step took 3.9 ms with the stock kernels, whose column reduction is a separate pass).
#include "common.cuh"
#include "det_reduce.cuh"

namespace {

constexpr int LN_WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VEC>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gamma,
                                                               const float4* __restrict__ beta, int rows, float eps,
                                                               float4* __restrict__ y, float* __restrict__ mean,
                                                               float* __restrict__ rstd) {
  constexpr int C4 = VEC * 32;                      // float4 per row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_n = 1.0f / (float)(C4 * 4);
  for (int r = blockIdx.x * LN_WARPS + warp; r < rows; r += gridDim.x * LN_WARPS) {
    float4 v[VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      v[i] = x[(size_t)r * C4 + i * 32 + lane];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mu = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rs = rsqrtf(warp_sum(q) * inv_n + eps);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 g = __ldg(gamma + i * 32 + lane), b = __ldg(beta + i * 32 + lane);
      y[(size_t)r * C4 + i * 32 + lane] =
          make_float4(v[i].x * rs * g.x + b.x, v[i].y * rs * g.y + b.y, v[i].z * rs * g.z + b.z, v[i].w * rs * g.w + b.w);
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

template <int VEC>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float4* __restrict__ gamma, int rows,
                                                               float4* __restrict__ dx, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, float* __restrict__ part,
                                                               unsigned* __restrict__ ticket) {
  constexpr int C4 = VEC * 32;
  __shared__ float4 red[LN_WARPS][C4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_n = 1.0f / (float)(C4 * 4);
  float4 gm[VEC], ag[VEC], ab[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    gm[i] = __ldg(gamma + i * 32 + lane);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = ag[i];
  }
  // two rows per iteration: the 4 * VEC loads of both rows are issued before either is consumed (the kernel streams 3 x 8 MB
  // per call and was latency bound with one row in flight per warp)
  auto row = [&](int r, const float4 (&xv)[VEC], const float4 (&d)[VEC]) {
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float4 xh[VEC], g[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      xh[i] = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);
      g[i] = make_float4(d[i].x * gm[i].x, d[i].y * gm[i].y, d[i].z * gm[i].z, d[i].w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      ag[i].x += d[i].x * xh[i].x; ag[i].y += d[i].y * xh[i].y; ag[i].z += d[i].z * xh[i].z; ag[i].w += d[i].w * xh[i].w;
      ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
    }
    const float m1 = warp_sum(s1) * inv_n, m2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      dx[(size_t)r * C4 + i * 32 + lane] =
          make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2), rs * (g[i].z - m1 - xh[i].z * m2),
                      rs * (g[i].w - m1 - xh[i].w * m2));
  };
  const int stride = gridDim.x * LN_WARPS;
  int r = blockIdx.x * LN_WARPS + warp;
  for (; r + stride < rows; r += 2 * stride) {
    float4 xa[VEC], da[VEC], xb[VEC], db[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      xa[i] = x[(size_t)r * C4 + i * 32 + lane]; da[i] = dy[(size_t)r * C4 + i * 32 + lane];
      xb[i] = x[(size_t)(r + stride) * C4 + i * 32 + lane]; db[i] = dy[(size_t)(r + stride) * C4 + i * 32 + lane];
    }
    row(r, xa, da);
    row(r + stride, xb, db);
  }
  if (r < rows) {
    float4 xa[VEC], da[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { xa[i] = x[(size_t)r * C4 + i * 32 + lane]; da[i] = dy[(size_t)r * C4 + i * 32 + lane]; }
    row(r, xa, da);
  }
  // column sums of this CTA: warps -> shared memory -> per-CTA partial sums -> last CTA adds them in block order
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[warp][i * 32 + lane] = pass == 0 ? ag[i] : ab[i];
    __syncthreads();
    for (int c = threadIdx.x; c < C4 * 4; c += LN_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) s += reinterpret_cast<const float*>(&red[w][0])[c];
      part[(size_t)blockIdx.x * 2 * (C4 * 4) + pass * (C4 * 4) + c] = s;
    }
    __syncthreads();
  }
  det_finish_columns<LN_WARPS * 32>(part, gridDim.x, 2 * C4 * 4, ticket, dgamma, dbeta, C4 * 4);
}

template <int VEC>
int launch_fwd(const float* x, const float* gamma, const float* beta, int rows, float eps, float* y, float* mean, float* rstd,
               cudaStream_t st) {
  const int grid = min((rows + LN_WARPS - 1) / LN_WARPS, vdetr_num_sms() * 8);
  ln_fwd_kernel<VEC><<<grid, LN_WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gamma),
                                                    reinterpret_cast<const float4*>(beta), rows, eps,
                                                    reinterpret_cast<float4*>(y), mean, rstd);
  VDETR_LAUNCH_CHECK();
  return 0;
}
template <int VEC>
int launch_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int rows, float* dx,
               float* dgamma, float* dbeta, float* ws, cudaStream_t st) {
  // enough rows per warp to amortise the column reduction; at most 64 CTAs (= partial sums of the final reduction)
  constexpr int C = VEC * 128;
  int grid = (rows + LN_WARPS * 4 - 1) / (LN_WARPS * 4);
  grid = max(1, min(grid, VDETR_RED_MAX_BLOCKS));          // (the order of the row sums depends on the grid only: deterministic)
  float* part = ws + 2 * C;
  unsigned* ticket = reinterpret_cast<unsigned*>(ws + (size_t)(VDETR_RED_MAX_BLOCKS + 1) * 2 * C);
  VDETR_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  ln_bwd_kernel<VEC><<<grid, LN_WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), mean,
                                                    rstd, reinterpret_cast<const float4*>(gamma), rows,
                                                    reinterpret_cast<float4*>(dx), dgamma, dbeta, part, ticket);
  VDETR_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int vdetr_layernorm_supported(int cols) { return cols == 128 || cols == 256 || cols == 384 || cols == 512; }

int vdetr_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, float eps, float* y,
                        float* mean, float* rstd, void* stream) {
  if (rows < 0 || !vdetr_layernorm_supported(cols)) return VDETR_ERR_UNSUPPORTED;
  if (rows == 0) return 0;
  if (!x || !gamma || !beta || !y || !mean || !rstd) return VDETR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (cols / 128) {
    case 1: return launch_fwd<1>(x, gamma, beta, rows, eps, y, mean, rstd, st);
    case 2: return launch_fwd<2>(x, gamma, beta, rows, eps, y, mean, rstd, st);
    case 3: return launch_fwd<3>(x, gamma, beta, rows, eps, y, mean, rstd, st);
    default: return launch_fwd<4>(x, gamma, beta, rows, eps, y, mean, rstd, st);
  }
}

int vdetr_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int rows,
                        int cols, float* dx, float* dgamma, float* dbeta, float* workspace, void* stream) {
  if (rows < 0 || !vdetr_layernorm_supported(cols)) return VDETR_ERR_UNSUPPORTED;
  if (!dgamma || !dbeta) return VDETR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    VDETR_CUDA_TRY(cudaMemsetAsync(dgamma, 0, (size_t)cols * sizeof(float), st));
    VDETR_CUDA_TRY(cudaMemsetAsync(dbeta, 0, (size_t)cols * sizeof(float), st));
    return 0;
  }
  if (!dy || !x || !mean || !rstd || !gamma || !dx || !workspace) return VDETR_ERR_BAD_ARG;
  switch (cols / 128) {
    case 1: return launch_bwd<1>(dy, x, mean, rstd, gamma, rows, dx, dgamma, dbeta, workspace, st);
    case 2: return launch_bwd<2>(dy, x, mean, rstd, gamma, rows, dx, dgamma, dbeta, workspace, st);
    case 3: return launch_bwd<3>(dy, x, mean, rstd, gamma, rows, dx, dgamma, dbeta, workspace, st);
    default: return launch_bwd<4>(dy, x, mean, rstd, gamma, rows, dx, dgamma, dbeta, workspace, st);
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Column sums of a dense [rows, cols] f32 matrix: the bias gradient of every token-major Linear / Conv1d(k=1)
// layer of the decoder (155 per step; the stock reduction kernel is latency bound on these 8 MB inputs).
namespace {
// grid = (row slices <= 64, column tiles of blockDim.x); a CTA sums its row slice for its columns, stores the partial sums,
// and the last CTA of the column tile (one ticket per tile) adds the slices in order: deterministic, no float atomics.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int rows, int cols, float* __restrict__ out,
                                                     float* __restrict__ part, unsigned* __restrict__ tickets) {
  __shared__ bool s_last;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  if (c < cols) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = r0;
    for (; r + 3 < r1; r += 4) {
      s0 += x[(size_t)r * cols + c]; s1 += x[(size_t)(r + 1) * cols + c];
      s2 += x[(size_t)(r + 2) * cols + c]; s3 += x[(size_t)(r + 3) * cols + c];
    }
    for (; r < r1; ++r) s0 += x[(size_t)r * cols + c];
    part[(size_t)blockIdx.x * cols + c] = (s0 + s1) + (s2 + s3);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(tickets + blockIdx.y, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last || c >= cols) return;
  __threadfence();
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = 0; b < (int)gridDim.x; ++b) s[b & 3] += __ldcg(part + (size_t)b * cols + c);
  out[c] = (s[0] + s[1]) + (s[2] + s[3]);
}
// cols % 4 == 0, cols <= 1024, 16-byte aligned rows: a thread owns 4 adjacent columns (one float4 per row), the cols / 4 threads
// of a row group walk the slice 8 rows at a time with all 8 loads in flight; row groups are combined through shared memory.
__global__ void __launch_bounds__(256) colsum4_kernel(const float4* __restrict__ x, int rows, int c4, float* __restrict__ out,
                                                      float* __restrict__ part, unsigned* __restrict__ tickets) {
  __shared__ float4 red[256];
  __shared__ bool s_last;
  const int groups = 256 / c4, grp = threadIdx.x / c4, col = threadIdx.x % c4;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  float4 acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (grp < groups) {
    int r = r0 + grp;
    for (; r + 7 * groups < r1; r += 8 * groups) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = x[(size_t)(r + j * groups) * c4 + col];
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j & 3].x += v[j].x; acc[j & 3].y += v[j].y; acc[j & 3].z += v[j].z; acc[j & 3].w += v[j].w; }
    }
    for (; r < r1; r += groups) {
      const float4 v = x[(size_t)r * c4 + col];
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
    }
  }
  red[threadIdx.x] = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                                 (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  __syncthreads();
  if (threadIdx.x < c4) {
    float4 s = red[threadIdx.x];
    for (int g = 1; g < groups; ++g) {
      const float4 t = red[g * c4 + threadIdx.x];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4*>(part)[(size_t)blockIdx.x * c4 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(tickets, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < c4 * 4; c += 256) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    int b = 0;
    for (; b + 7 < (int)gridDim.x; b += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += __ldcg(part + (size_t)(b + j) * c4 * 4 + c);
    }
    for (int j = 0; b < (int)gridDim.x; ++b, ++j) s[j] += __ldcg(part + (size_t)b * c4 * 4 + c);
    out[c] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  }
}
}  // namespace

// workspace: vdetr_colsum_workspace_floats(cols) floats
extern "C" size_t vdetr_colsum_workspace_floats(int cols) { return cols > 0 ? (size_t)VDETR_RED_MAX_BLOCKS * cols + 64 : 0; }

extern "C" int vdetr_colsum(const float* x, int rows, int cols, float* out, float* workspace, void* stream) {
  if (rows < 0 || cols < 1) return VDETR_ERR_BAD_ARG;
  if (!out || (rows > 0 && (!x || !workspace))) return VDETR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    VDETR_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), st));
    return 0;
  }
  if (cols % 4 == 0 && cols <= 1024 && 256 % (cols / 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && rows >= 256) {
    int slices4 = (rows + 127) / 128;
    slices4 = slices4 > VDETR_RED_MAX_BLOCKS ? VDETR_RED_MAX_BLOCKS : slices4;
    unsigned* tk = reinterpret_cast<unsigned*>(workspace + (size_t)VDETR_RED_MAX_BLOCKS * cols);
    VDETR_CUDA_TRY(cudaMemsetAsync(tk, 0, sizeof(unsigned), st));
    colsum4_kernel<<<slices4, 256, 0, st>>>(reinterpret_cast<const float4*>(x), rows, cols / 4, out, workspace, tk);
    VDETR_LAUNCH_CHECK();
    return 0;
  }
  const int threads = cols >= 256 ? 256 : ((cols + 31) / 32) * 32;
  const int tiles = (cols + threads - 1) / threads;
  if (tiles > 64) return VDETR_ERR_UNSUPPORTED;
  int slices = (rows + 63) / 64;
  slices = slices > VDETR_RED_MAX_BLOCKS ? VDETR_RED_MAX_BLOCKS : slices;
  unsigned* tickets = reinterpret_cast<unsigned*>(workspace + (size_t)VDETR_RED_MAX_BLOCKS * cols);
  VDETR_CUDA_TRY(cudaMemsetAsync(tickets, 0, 64 * sizeof(unsigned), st));
  dim3 grid(slices, tiles);
  colsum_kernel<<<grid, threads, 0, st>>>(x, rows, cols, out, workspace, tickets);
  VDETR_LAUNCH_CHECK();
  return 0;
}
