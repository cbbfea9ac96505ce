// Training-mode BatchNorm1d + ReLU on token-major activations [tokens, C] (the Conv1d(k=1)-BN-ReLU stacks of the box
// heads and of the query-position MLP: models/helpers.py:17-33, 74-141, evaluated token-major -- helpers.pointwise_tokens).
//   stats   : per-channel sum and sum of squares of (x - pivot), pivot = row 0 (shifted sums: no cancellation when
//             |mean| >> std); warps accumulate per-lane column sums over their rows, CTAs reduce in shared memory, store
//             their partial sums, and the last CTA to finish adds the partials in block order (det_reduce.cuh: no float
//             atomics, so two identical calls give identical bits)
//   apply   : y = relu((x - mean) * rstd * gamma + beta); block 0 also updates running_mean / running_var (unbiased)
//   bwd sums: dbeta = sum g, dgamma = sum g * xhat with g = dy * [y > 0]  (the two column sums the input gradient needs)
//   bwd dx  : dx = gamma * rstd * (g - dbeta / T - xhat * dgamma / T)
// Stock PyTorch runs 3 kernels forward (statistics, transform, ReLU) and 3 backward; here the ReLU rides along -- and so
// does the nn.Dropout that follows it in the reference's stacks (models/helpers.py:118-120, --mlp_dropout 0.3): the apply
// kernel multiplies by keep / (1 - p) (Philox mask, philox.cuh), and because a dropped or clipped element is stored as 0 the
// backward needs no mask at all: g = dy * [y > 0] / (1 - p).
// Several independent BatchNorm layers of the same width are normalised by ONE launch (the 5 box heads of a decoder level
// evaluated together): blockIdx.y = group; element (row r, group g, channel c) is x[g * group_stride + r * row_stride + c],
// which covers both the channels-last [tokens, G * C] output of a concatenated GEMM and the head-major [G, tokens, C]
// output of a batched one.  gs4 / ld4 below are group_stride / 4 and row_stride / 4.
#include "common.cuh"
#include "det_reduce.cuh"
#include "philox.cuh"

namespace {

constexpr int BN_WARPS = 8;

// column accumulation helper: every lane owns VEC float4 (columns i*128 + lane*4 .. +3).  The CTA's sums of the two
// quantities go to part[blockIdx.x][2 * C]; the last CTA writes the totals to out_a[C] / out_b[C].
template <int VEC>
__device__ __forceinline__ void reduce_columns_to_global(float4 (&a)[VEC], float4 (&b)[VEC], float* out_a, float* out_b, float* part,
                                                         unsigned* ticket) {
  __shared__ float4 red[BN_WARPS][VEC * 32];
  constexpr int C = VEC * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[warp][i * 32 + lane] = pass == 0 ? a[i] : b[i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += BN_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < BN_WARPS; ++w) s += reinterpret_cast<const float*>(&red[w][0])[c];
      part[(size_t)blockIdx.x * 2 * C + pass * C + c] = s;
    }
    __syncthreads();
  }
  det_finish_columns<BN_WARPS * 32>(part, gridDim.x, 2 * C, ticket, out_a, out_b, C);
}
// per-chunk views of the reduction workspace: partial sums [chunk][64][2 * C], one ticket per chunk
template <int VEC>
__device__ __forceinline__ float* chunk_part(float* part) { return part + (size_t)blockIdx.y * VDETR_RED_MAX_BLOCKS * 2 * (VEC * 128); }

template <int VEC>
__global__ void __launch_bounds__(BN_WARPS * 32) bn_stats_kernel(const float4* __restrict__ x, int rows, float* __restrict__ sum,
                                                                 float* __restrict__ sumsq, float* __restrict__ part,
                                                                 unsigned* __restrict__ ticket, int ld4, size_t gs4) {
  constexpr int C = VEC * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  x += blockIdx.y * gs4; sum += blockIdx.y * C; sumsq += blockIdx.y * C;
  part = chunk_part<VEC>(part); ticket += blockIdx.y;
  float4 s[VEC], q[VEC], pv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    q[i] = s[i];
    pv[i] = __ldg(x + i * 32 + lane);                   // pivot = row 0
  }
  for (int r = blockIdx.x * BN_WARPS + warp; r < rows; r += gridDim.x * BN_WARPS) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 v = x[(size_t)r * ld4 + i * 32 + lane];
      v.x -= pv[i].x; v.y -= pv[i].y; v.z -= pv[i].z; v.w -= pv[i].w;
      s[i].x += v.x; s[i].y += v.y; s[i].z += v.z; s[i].w += v.w;
      q[i].x += v.x * v.x; q[i].y += v.y * v.y; q[i].z += v.z * v.z; q[i].w += v.w * v.w;
    }
  }
  reduce_columns_to_global<VEC>(s, q, sum, sumsq, part, ticket);
}

template <int VEC>
__global__ void __launch_bounds__(BN_WARPS * 32) bn_apply_relu_kernel(const float4* __restrict__ x, int rows,
                                                                      const float* __restrict__ sum, const float* __restrict__ sumsq,
                                                                      const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                                                      float eps, float momentum, float4* __restrict__ y,
                                                                      float* __restrict__ mean, float* __restrict__ rstd,
                                                                      float* running_mean, float* running_var, int ld4, size_t gs4,
                                                                      uint32_t drop_thresh, float inv_keep,
                                                                      const unsigned long long* __restrict__ drop_seed,
                                                                      const float* __restrict__ total_rows) {
  constexpr int C4 = VEC * 32, C = C4 * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2 dkey = make_uint2(0u, 0u);
  if (drop_thresh) {
    const unsigned long long sd = __ldg(drop_seed);
    dkey = make_uint2((uint32_t)sd, (uint32_t)(sd >> 32));
  }
  x += blockIdx.y * gs4; y += blockIdx.y * gs4; gamma += blockIdx.y * C4; beta += blockIdx.y * C4;
  sum += blockIdx.y * C; sumsq += blockIdx.y * C; mean += blockIdx.y * C; rstd += blockIdx.y * C;
  if (running_mean) { running_mean += blockIdx.y * C; running_var += blockIdx.y * C; }
  const float inv_n = 1.0f / (float)rows;
  float4 sc[VEC], sh[VEC];            // y = relu(x * sc + sh)
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c0 = (i * 32 + lane) * 4;
    const float4 pv = __ldg(x + i * 32 + lane), g = __ldg(gamma + i * 32 + lane), b = __ldg(beta + i * 32 + lane);
    float m[4], rs[4];
    const float pvv[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ms = __ldg(sum + c0 + e) * inv_n;                       // mean of (x - pivot)
      const float var = fmaxf(__ldg(sumsq + c0 + e) * inv_n - ms * ms, 0.f);
      m[e] = ms + pvv[e];
      rs[e] = rsqrtf(var + eps);
      if (blockIdx.x == 0 && warp == 0) {
        mean[c0 + e] = m[e];
        rstd[c0 + e] = rs[e];
        if (running_mean) {
          const float nt = total_rows ? __ldg(total_rows) : (float)rows;      // SyncBatchNorm: rows of all ranks
          const float unbiased = nt > 1.f ? var * (nt / (nt - 1.f)) : var;
          running_mean[c0 + e] = (1.f - momentum) * running_mean[c0 + e] + momentum * m[e];
          running_var[c0 + e] = (1.f - momentum) * running_var[c0 + e] + momentum * unbiased;
        }
      }
    }
    sc[i] = make_float4(rs[0] * g.x, rs[1] * g.y, rs[2] * g.z, rs[3] * g.w);
    sh[i] = make_float4(b.x - m[0] * sc[i].x, b.y - m[1] * sc[i].y, b.z - m[2] * sc[i].z, b.w - m[3] * sc[i].w);
  }
  (void)C;
  for (int r = blockIdx.x * BN_WARPS + warp; r < rows; r += gridDim.x * BN_WARPS) {
    uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 v = x[(size_t)r * ld4 + i * 32 + lane];
      float4 o = make_float4(fmaxf(fmaf(v.x, sc[i].x, sh[i].x), 0.f), fmaxf(fmaf(v.y, sc[i].y, sh[i].y), 0.f),
                             fmaxf(fmaf(v.z, sc[i].z, sh[i].z), 0.f), fmaxf(fmaf(v.w, sc[i].w, sh[i].w), 0.f));
      if (drop_thresh) {
        // one Philox call = eight 16-bit numbers = the two float4 (i even / odd) of this lane; counter = (row, group, column pair)
        if ((i & 1) == 0) rnd = philox::philox4x32_10(make_uint4((uint32_t)r, blockIdx.y, (uint32_t)((i >> 1) * 32 + lane), 0x42u), dkey);
        const uint32_t w0 = (i & 1) ? rnd.z : rnd.x, w1 = (i & 1) ? rnd.w : rnd.y;
        o.x = (w0 & 0xFFFFu) >= drop_thresh ? o.x * inv_keep : 0.f;
        o.y = (w0 >> 16) >= drop_thresh ? o.y * inv_keep : 0.f;
        o.z = (w1 & 0xFFFFu) >= drop_thresh ? o.z * inv_keep : 0.f;
        o.w = (w1 >> 16) >= drop_thresh ? o.w * inv_keep : 0.f;
      }
      y[(size_t)r * ld4 + i * 32 + lane] = o;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(BN_WARPS * 32) bn_bwd_sums_kernel(const float4* __restrict__ dy, const float4* __restrict__ y,
                                                                    const float4* __restrict__ x, int rows,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                    float* __restrict__ part, unsigned* __restrict__ ticket, int ld4,
                                                                    size_t gs4, float inv_keep) {
  constexpr int C = VEC * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  dy += blockIdx.y * gs4; y += blockIdx.y * gs4; x += blockIdx.y * gs4;
  mean += blockIdx.y * C; rstd += blockIdx.y * C; dgamma += blockIdx.y * C; dbeta += blockIdx.y * C;
  part = chunk_part<VEC>(part); ticket += blockIdx.y;
  float4 ag[VEC], ab[VEC], mu[VEC], rs[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = ag[i];
    mu[i] = __ldg(reinterpret_cast<const float4*>(mean) + i * 32 + lane);
    rs[i] = __ldg(reinterpret_cast<const float4*>(rstd) + i * 32 + lane);
  }
  for (int r = blockIdx.x * BN_WARPS + warp; r < rows; r += gridDim.x * BN_WARPS) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const size_t o = (size_t)r * ld4 + i * 32 + lane;
      const float4 d = dy[o], yy = y[o], xv = x[o];
      const float gx = yy.x > 0.f ? d.x * inv_keep : 0.f, gy = yy.y > 0.f ? d.y * inv_keep : 0.f;
      const float gz = yy.z > 0.f ? d.z * inv_keep : 0.f, gw = yy.w > 0.f ? d.w * inv_keep : 0.f;
      ab[i].x += gx; ab[i].y += gy; ab[i].z += gz; ab[i].w += gw;
      ag[i].x += gx * (xv.x - mu[i].x) * rs[i].x; ag[i].y += gy * (xv.y - mu[i].y) * rs[i].y;
      ag[i].z += gz * (xv.z - mu[i].z) * rs[i].z; ag[i].w += gw * (xv.w - mu[i].w) * rs[i].w;
    }
  }
  reduce_columns_to_global<VEC>(ag, ab, dgamma, dbeta, part, ticket);
}

template <int VEC>
__global__ void __launch_bounds__(BN_WARPS * 32) bn_bwd_dx_kernel(const float4* __restrict__ dy, const float4* __restrict__ y,
                                                                  const float4* __restrict__ x, int rows,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float4* __restrict__ gamma, const float* __restrict__ dgamma,
                                                                  const float* __restrict__ dbeta, float4* __restrict__ dx, int ld4,
                                                                  size_t gs4, float inv_keep, const float* __restrict__ total_rows) {
  constexpr int C4 = VEC * 32, C = VEC * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  dy += blockIdx.y * gs4; y += blockIdx.y * gs4; x += blockIdx.y * gs4; dx += blockIdx.y * gs4; gamma += blockIdx.y * C4;
  mean += blockIdx.y * C; rstd += blockIdx.y * C; dgamma += blockIdx.y * C; dbeta += blockIdx.y * C;
  const float inv_n = 1.0f / (total_rows ? __ldg(total_rows) : (float)rows);      // SyncBatchNorm: sums and rows of all ranks
  float4 mu[VEC], rs[VEC], k1[VEC], k2[VEC], k3[VEC];      // dx = k1 * g - k2 - xhat * k3
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    mu[i] = __ldg(reinterpret_cast<const float4*>(mean) + i * 32 + lane);
    rs[i] = __ldg(reinterpret_cast<const float4*>(rstd) + i * 32 + lane);
    const float4 g = __ldg(gamma + i * 32 + lane);
    const float4 dg = __ldg(reinterpret_cast<const float4*>(dgamma) + i * 32 + lane);
    const float4 db = __ldg(reinterpret_cast<const float4*>(dbeta) + i * 32 + lane);
    k1[i] = make_float4(g.x * rs[i].x, g.y * rs[i].y, g.z * rs[i].z, g.w * rs[i].w);
    k2[i] = make_float4(k1[i].x * db.x * inv_n, k1[i].y * db.y * inv_n, k1[i].z * db.z * inv_n, k1[i].w * db.w * inv_n);
    k3[i] = make_float4(k1[i].x * dg.x * inv_n, k1[i].y * dg.y * inv_n, k1[i].z * dg.z * inv_n, k1[i].w * dg.w * inv_n);
  }
  for (int r = blockIdx.x * BN_WARPS + warp; r < rows; r += gridDim.x * BN_WARPS) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const size_t o = (size_t)r * ld4 + i * 32 + lane;
      const float4 d = dy[o], yy = y[o], xv = x[o];
      const float gx = yy.x > 0.f ? d.x * inv_keep : 0.f, gy = yy.y > 0.f ? d.y * inv_keep : 0.f;
      const float gz = yy.z > 0.f ? d.z * inv_keep : 0.f, gw = yy.w > 0.f ? d.w * inv_keep : 0.f;
      dx[o] = make_float4(k1[i].x * gx - k2[i].x - (xv.x - mu[i].x) * rs[i].x * k3[i].x,
                          k1[i].y * gy - k2[i].y - (xv.y - mu[i].y) * rs[i].y * k3[i].y,
                          k1[i].z * gz - k2[i].z - (xv.z - mu[i].z) * rs[i].z * k3[i].z,
                          k1[i].w * gw - k2[i].w - (xv.w - mu[i].w) * rs[i].w * k3[i].w);
    }
  }
}

// SyncBatchNorm context (vdetr_bn_sync_set): world > 1 makes every training BatchNorm of this process exchange its statistics
// with the other ranks -- the library-wide equivalent of nn.SyncBatchNorm.convert_sync_batchnorm(model), main.py:512-514.
VdetrPeerCtx g_bn_sync = {};

inline int bn_grid(int rows, int rows_per_warp) {
  int g = (rows + BN_WARPS * rows_per_warp - 1) / (BN_WARPS * rows_per_warp);
  return g < 1 ? 1 : (g > vdetr_num_sms() * 4 ? vdetr_num_sms() * 4 : g);
}
inline int bn_red_grid(int rows) {                 // kernels that end in a cross-CTA reduction: at most 64 partial sums
  const int g = bn_grid(rows, 4);
  return g > VDETR_RED_MAX_BLOCKS ? VDETR_RED_MAX_BLOCKS : g;
}

struct BnLayout {
  int groups;
  size_t group_stride, row_stride;      // in floats
};

template <int VEC>
int fwd_t(const float* x, const float* gamma, const float* beta, int rows, const BnLayout& L, float eps, float momentum, float* y,
          float* mean, float* rstd, float* running_mean, float* running_var, float* ws, float drop_p,
          const unsigned long long* drop_seed, cudaStream_t st) {
  const int cols = VEC * 128 * L.groups, ld4 = (int)(L.row_stride / 4);
  const size_t gs4 = L.group_stride / 4;
  float* part = ws + 2 * cols;
  unsigned* tickets = reinterpret_cast<unsigned*>(ws + (size_t)(VDETR_RED_MAX_BLOCKS + 1) * 2 * cols);
  VDETR_CUDA_TRY(cudaMemsetAsync(tickets, 0, 32 * sizeof(unsigned), st));
  bn_stats_kernel<VEC><<<dim3(bn_red_grid(rows), L.groups), BN_WARPS * 32, 0, st>>>(reinterpret_cast<const float4*>(x), rows, ws,
                                                                                    ws + cols, part, tickets, ld4, gs4);
  VDETR_LAUNCH_CHECK();
  float* misc = ws + (size_t)(VDETR_RED_MAX_BLOCKS + 1) * 2 * cols + 32;       // [0]: rows of all ranks
  const float* total_rows = nullptr;
  if (g_bn_sync.world > 1) {      // SyncBatchNorm: merge the statistics of all ranks over peer memory (peer.cu)
    const int rc = vdetr_peer_bn_fwd(g_bn_sync.flags, g_bn_sync.slots, g_bn_sync.rank, g_bn_sync.world, g_bn_sync.epoch, g_bn_sync.cap,
                                     ws, ws + cols, x, VEC * 128, L.groups, (long long)L.group_stride, rows, misc, st);
    if (rc) return rc;
    total_rows = misc;
  }
  bn_apply_relu_kernel<VEC><<<dim3(bn_grid(rows, 2), L.groups), BN_WARPS * 32, 0, st>>>(
      reinterpret_cast<const float4*>(x), rows, ws, ws + cols, reinterpret_cast<const float4*>(gamma),
      reinterpret_cast<const float4*>(beta), eps, momentum, reinterpret_cast<float4*>(y), mean, rstd, running_mean, running_var, ld4, gs4,
      drop_p > 0.f ? philox::thresh_of(drop_p) : 0u, 1.0f / (1.0f - drop_p), drop_seed, total_rows);
  VDETR_LAUNCH_CHECK();
  return 0;
}
template <int VEC>
int bwd_t(const float* dy, const float* y, const float* x, const float* mean, const float* rstd, const float* gamma, int rows,
          const BnLayout& L, float* dx, float* dgamma, float* dbeta, float* ws, float drop_p, cudaStream_t st) {
  const float inv_keep = 1.0f / (1.0f - drop_p);
  const int cols = VEC * 128 * L.groups, ld4 = (int)(L.row_stride / 4);
  const size_t gs4 = L.group_stride / 4;
  float* part = ws + 2 * cols;
  unsigned* tickets = reinterpret_cast<unsigned*>(ws + (size_t)(VDETR_RED_MAX_BLOCKS + 1) * 2 * cols);
  VDETR_CUDA_TRY(cudaMemsetAsync(tickets, 0, 32 * sizeof(unsigned), st));
  bn_bwd_sums_kernel<VEC><<<dim3(bn_red_grid(rows), L.groups), BN_WARPS * 32, 0, st>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(x), rows, mean, rstd,
      dgamma, dbeta, part, tickets, ld4, gs4, inv_keep);
  VDETR_LAUNCH_CHECK();
  float* misc = ws + (size_t)(VDETR_RED_MAX_BLOCKS + 1) * 2 * cols + 32;       // [0]: rows of all ranks, [32 ...): sums of all ranks
  const float *sum_g = dgamma, *sum_b = dbeta, *total_rows = nullptr;
  if (g_bn_sync.world > 1) {      // SyncBatchNorm: dx needs the two sums over ALL ranks; dgamma / dbeta stay this rank's
    const int rc = vdetr_peer_bn_bwd(g_bn_sync.flags, g_bn_sync.slots, g_bn_sync.rank, g_bn_sync.world, g_bn_sync.epoch, g_bn_sync.cap,
                                     dgamma, dbeta, cols, rows, misc + 32, misc + 32 + cols, misc, st);
    if (rc) return rc;
    sum_g = misc + 32; sum_b = misc + 32 + cols; total_rows = misc;
  }
  bn_bwd_dx_kernel<VEC><<<dim3(bn_grid(rows, 2), L.groups), BN_WARPS * 32, 0, st>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(x), rows, mean, rstd,
      reinterpret_cast<const float4*>(gamma), sum_g, sum_b, reinterpret_cast<float4*>(dx), ld4, gs4, inv_keep, total_rows);
  VDETR_LAUNCH_CHECK();
  return 0;
}

bool layout_ok(int cols, int groups, long long group_stride, long long row_stride) {
  return groups >= 1 && groups <= 32 && group_stride % 4 == 0 && row_stride % 4 == 0 && row_stride >= cols &&
         (groups == 1 || group_stride >= cols);
}

}  // namespace

extern "C" {

int vdetr_bn_sync_set(const VdetrPeerCtx* ctx) {
  if (!ctx) { g_bn_sync = VdetrPeerCtx{}; return 0; }
  if (ctx->world < 1 || ctx->world > VDETR_PEER_MAX_WORLD || ctx->rank < 0 || ctx->rank >= ctx->world || !ctx->epoch || ctx->cap < 1)
    return VDETR_ERR_BAD_ARG;
  g_bn_sync = *ctx;
  return 0;
}

int vdetr_bn_relu_supported(int cols) { return cols == 128 || cols == 256 || cols == 384 || cols == 512; }

size_t vdetr_reduce_workspace_floats(int cols) { return cols > 0 ? vdetr_reduce_ws_floats(cols) : 0; }

int vdetr_bn_relu_train_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, int groups,
                            long long group_stride, long long row_stride, float eps, float momentum, float dropout_p,
                            const uint64_t* dropout_seed, float* y, float* mean, float* rstd, float* running_mean,
                            float* running_var, float* workspace, void* stream) {
  if (rows < 1 || !vdetr_bn_relu_supported(cols) || !layout_ok(cols, groups, group_stride, row_stride)) return VDETR_ERR_UNSUPPORTED;
  if (!x || !gamma || !beta || !y || !mean || !rstd || !workspace) return VDETR_ERR_BAD_ARG;
  if (!(dropout_p >= 0.f) || dropout_p >= 1.f || (dropout_p > 0.f && !dropout_seed)) return VDETR_ERR_BAD_ARG;
  const unsigned long long* seed = reinterpret_cast<const unsigned long long*>(dropout_seed);
  cudaStream_t st = (cudaStream_t)stream;
  const BnLayout L = {groups, (size_t)group_stride, (size_t)row_stride};
  switch (cols / 128) {
    case 1: return fwd_t<1>(x, gamma, beta, rows, L, eps, momentum, y, mean, rstd, running_mean, running_var, workspace, dropout_p, seed, st);
    case 2: return fwd_t<2>(x, gamma, beta, rows, L, eps, momentum, y, mean, rstd, running_mean, running_var, workspace, dropout_p, seed, st);
    case 3: return fwd_t<3>(x, gamma, beta, rows, L, eps, momentum, y, mean, rstd, running_mean, running_var, workspace, dropout_p, seed, st);
    default: return fwd_t<4>(x, gamma, beta, rows, L, eps, momentum, y, mean, rstd, running_mean, running_var, workspace, dropout_p, seed, st);
  }
}

int vdetr_bn_relu_train_bwd(const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
                            const float* gamma, int rows, int cols, int groups, long long group_stride, long long row_stride,
                            float dropout_p, float* dx, float* dgamma, float* dbeta, float* workspace, void* stream) {
  if (rows < 1 || !vdetr_bn_relu_supported(cols) || !layout_ok(cols, groups, group_stride, row_stride)) return VDETR_ERR_UNSUPPORTED;
  if (!(dropout_p >= 0.f) || dropout_p >= 1.f) return VDETR_ERR_BAD_ARG;
  if (!dy || !y || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || !workspace) return VDETR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const BnLayout L = {groups, (size_t)group_stride, (size_t)row_stride};
  switch (cols / 128) {
    case 1: return bwd_t<1>(dy, y, x, mean, rstd, gamma, rows, L, dx, dgamma, dbeta, workspace, dropout_p, st);
    case 2: return bwd_t<2>(dy, y, x, mean, rstd, gamma, rows, L, dx, dgamma, dbeta, workspace, dropout_p, st);
    case 3: return bwd_t<3>(dy, y, x, mean, rstd, gamma, rows, L, dx, dgamma, dbeta, workspace, dropout_p, st);
    default: return bwd_t<4>(dy, y, x, mean, rstd, gamma, rows, L, dx, dgamma, dbeta, workspace, dropout_p, st);
  }
}

}  // extern "C"
