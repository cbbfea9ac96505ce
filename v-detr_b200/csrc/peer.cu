// Peer memory over NVLink / NVSwitch for the data-parallel step (one process per GPU), and the two places where the step
// has a real exchange:
//
//   * gradient all-reduce + optimizer (the reference: DistributedDataParallel's bucketed NCCL all-reduce, main.py:515-517,
//     then clip_grad_norm_ + AdamW, engine.py:105-108) as ONE pass over peer memory: every rank owns a contiguous shard
//     of the flat parameter vector; it PULLS the gradients of its shard from all ranks (P2P loads, summed in rank order, so
//     every rank of every run sees the same bits), takes its share of the squared gradient norm, and -- after a flag
//     barrier that also exchanges the W x G partial norms -- applies AdamW to its shard and PUSHES the new parameters into
//     every rank's copy (P2P stores).  Reduce-scatter -> update -> all-gather, 2 x (W-1)/W x 46.5 MB over NVLink per GPU and
//     step instead of an NCCL all-reduce followed by a full-size optimizer pass on every GPU; the moments exist only for
//     the shard (ZeRO-1).
//   * SyncBatchNorm statistics (main.py:512-514 converts every BatchNorm to SyncBatchNorm): the per-channel partial sums a
//     BatchNorm kernel has just reduced on its GPU are pushed into every peer's slot, a flag barrier follows, and every
//     rank merges the W partials in rank order (Chan's parallel variance) -- see vdetr_peer_bn_* below; ~50 exchanges of
//     2 KB per step, each one tiny kernel instead of the ~200 NCCL launches of torch.nn.SyncBatchNorm.
//
// Buffers come from cudaMalloc and travel between the processes as CUDA IPC handles (the Python side exchanges the 64-byte
// handles through torch.distributed).  Flags: flags[channel][rank] (32-bit epochs) in every rank's peer memory; a barrier
// on `channel` writes the next epoch into slot [my rank] of every peer with a system-scope release and spins on its own
// slots with system-scope acquires.  The epoch counters live in device memory, so a captured CUDA graph that contains
// barriers can be replayed: every replay advances them.  A spin gives up after VDETR_PEER_TIMEOUT_NS and raises a sticky
// device-side error flag (vdetr_peer_error) instead of hanging the GPU when a peer has died.
#include "common.cuh"
#include <string.h>

namespace {

constexpr int MAXW = VDETR_PEER_MAX_WORLD;
constexpr unsigned long long TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

__device__ int g_peer_error = 0;

struct Peers {
  void* p[MAXW];
};

__device__ __forceinline__ void st_release_sys(unsigned* addr, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* addr) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One CTA.  next_epoch: all threads get the epoch of the barrier this kernel is about to run (the counter is this rank's
// private device memory).  signal_wait: thread r < world signals peer r and waits for peer r.
__device__ __forceinline__ unsigned next_epoch(unsigned* epoch, int channel) {
  __shared__ unsigned s_epoch;
  __syncthreads();
  if (threadIdx.x == 0) s_epoch = ++epoch[channel];
  __syncthreads();
  return s_epoch;
}
__device__ __forceinline__ void signal_wait(const Peers& flags, int rank, int world, int channel, unsigned e) {
  __threadfence_system();                       // everything this GPU wrote before the barrier (other kernels included)
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int r = threadIdx.x;
    st_release_sys(reinterpret_cast<unsigned*>(flags.p[r]) + channel * MAXW + rank, e);
    const unsigned* mine = reinterpret_cast<const unsigned*>(flags.p[rank]) + channel * MAXW + r;
    const unsigned long long t0 = globaltimer();
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if (globaltimer() - t0 > TIMEOUT_NS) { g_peer_error = 1; break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void flag_barrier(const Peers& flags, int rank, int world, unsigned* epoch, int channel) {
  signal_wait(flags, rank, world, channel, next_epoch(epoch, channel));
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(Peers flags, int rank, int world, unsigned* epoch, int channel) {
  flag_barrier(flags, rank, world, epoch, channel);
}

// ------------------------------------------------------------------------------------------------ optimizer
struct PeerAdam {
  Peers p, g;                          // every rank's flat parameter / gradient vector (n floats)
  Peers norm_part;                     // every rank's [MAXW][NORM_BLOCKS] partial squared norms
  float* red;                          // local: the reduced gradient of the shard (shard floats)
  float* m; float* v;                  // local: moments of the shard
  long long n, n_decay, lo, hi;        // the shard is [lo, hi); lo and hi are multiples of 4 except hi == n
  const float* lr; const float* step;
  float* norm_out;                     // local [1]: global gradient norm (after the 1 / world scale)
  float gscale_host, max_norm, b1, b2, eps, wd;
  int rank, world;
};
constexpr int NORM_BLOCKS = 256;

__global__ void __launch_bounds__(256) peer_reduce_kernel(const PeerAdam A) {
  __shared__ float wsum[8];
  float acc = 0.f;
  const long long n4 = (A.hi - A.lo) / 4;
  const long long lo4 = A.lo / 4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < A.world; ++r) {          // rank order: identical bits on every rank
      const float4 g = reinterpret_cast<const float4*>(A.g.p[r])[lo4 + i];
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    reinterpret_cast<float4*>(A.red)[i] = s;
    acc += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
  }
  if (blockIdx.x == 0) {                          // tail of the last shard (n not a multiple of 4)
    const long long t = A.lo + n4 * 4 + threadIdx.x;
    if (t < A.hi) {
      float s = 0.f;
      for (int r = 0; r < A.world; ++r) s += reinterpret_cast<const float*>(A.g.p[r])[t];
      A.red[t - A.lo] = s;
      acc += s * s;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    for (int r = 0; r < A.world; ++r)            // push this CTA's share of the squared norm to every rank
      reinterpret_cast<float*>(A.norm_part.p[r])[A.rank * NORM_BLOCKS + blockIdx.x] = t;
  }
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float decay, float b1, float b2, float eps, float step_size,
                                      float inv_sqrt_bc2) {
  p *= decay;
  m = m + (1.f - b1) * (g - m);
  v = v * b2 + (1.f - b2) * g * g;
  p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
}

__global__ void __launch_bounds__(256) peer_adamw_kernel(const PeerAdam A) {
  // global squared norm: W x NORM_BLOCKS partials, added in a fixed order by every CTA (same bits everywhere)
  __shared__ float s_norm2;
  if (threadIdx.x < 32) {
    const float* part = reinterpret_cast<const float*>(A.norm_part.p[A.rank]);
    float t = 0.f;
    for (int i = threadIdx.x; i < A.world * NORM_BLOCKS; i += 32) t += part[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) s_norm2 = t;
  }
  __syncthreads();
  const float norm = sqrtf(s_norm2) * A.gscale_host;           // norm of the AVERAGED gradient (gscale_host = 1 / world)
  if (blockIdx.x == 0 && threadIdx.x == 0 && A.norm_out) *A.norm_out = norm;
  const float clip = A.max_norm > 0.f ? fminf(A.max_norm / (norm + 1e-6f), 1.f) : 1.f;
  const float gs = A.gscale_host * clip;
  const float lr = __ldg(A.lr), t = __ldg(A.step);
  const float bc1 = 1.f - powf(A.b1, t), bc2 = 1.f - powf(A.b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay_on = 1.f - lr * A.wd;
  const long long n4 = (A.hi - A.lo) / 4, lo4 = A.lo / 4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 p = reinterpret_cast<const float4*>(A.p.p[A.rank])[lo4 + i];
    float4 m = reinterpret_cast<float4*>(A.m)[i], v = reinterpret_cast<float4*>(A.v)[i];
    float4 g = reinterpret_cast<const float4*>(A.red)[i];
    g.x *= gs; g.y *= gs; g.z *= gs; g.w *= gs;
    const long long e = A.lo + i * 4;
    adam1(p.x, g.x, m.x, v.x, e + 0 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    adam1(p.y, g.y, m.y, v.y, e + 1 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    adam1(p.z, g.z, m.z, v.z, e + 2 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    adam1(p.w, g.w, m.w, v.w, e + 3 < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
    reinterpret_cast<float4*>(A.m)[i] = m; reinterpret_cast<float4*>(A.v)[i] = v;
    for (int r = 0; r < A.world; ++r) reinterpret_cast<float4*>(A.p.p[r])[lo4 + i] = p;     // all-gather by P2P stores
  }
  if (blockIdx.x == 0) {
    const long long tl = A.lo + n4 * 4 + threadIdx.x;
    if (tl < A.hi) {
      float p = reinterpret_cast<const float*>(A.p.p[A.rank])[tl], m = A.m[tl - A.lo], v = A.v[tl - A.lo];
      adam1(p, A.red[tl - A.lo] * gs, m, v, tl < A.n_decay ? decay_on : 1.f, A.b1, A.b2, A.eps, step_size, inv_sqrt_bc2);
      A.m[tl - A.lo] = m; A.v[tl - A.lo] = v;
      for (int r = 0; r < A.world; ++r) reinterpret_cast<float*>(A.p.p[r])[tl] = p;
    }
  }
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------------ SyncBatchNorm statistics
// Exchange buffer of every rank: [2 (epoch parity)][MAXW (source rank)][1 + 2 * cap] floats.  Consecutive exchanges use
// alternating halves: a rank can be at most one barrier ahead of a peer, so the half a fast rank writes for exchange k + 1
// is never the one a slow rank still reads for exchange k.
//
// Forward.  Input: this rank's shifted sums over its `rows` tokens (sum / sumsq of x - pivot, pivot = its row 0: what
// bn_stats_kernel leaves), for G groups of C channels (channel gc of group g: pivot = x[g * group_stride + c]).  Every rank
// pushes (rows, mean, M2) per channel to all peers, barrier, merge in rank order (Chan et al.), and rewrites sum / sumsq IN
// PLACE so that bn_apply_relu_kernel -- which reads "shifted sums over `rows` local rows" -- reproduces the GLOBAL mean and
// biased variance; total_rows_out gets the global row count (for the unbiased running variance).
__global__ void __launch_bounds__(1024) peer_bn_fwd_kernel(Peers flags, Peers slots, int rank, int world, unsigned* epoch, int channel,
                                                          int cap, float* sum, float* sumsq, const float* x, int C, int G,
                                                          long long group_stride, int rows, float* total_rows_out) {
  const int cols = C * G;
  const size_t per = 1 + 2 * (size_t)cap;
  const unsigned e = next_epoch(epoch, channel);
  const size_t half = (size_t)(e & 1u) * MAXW * per;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float pv = x[(size_t)(c / C) * group_stride + (c % C)];
    const float ms = sum[c] / (float)rows;
    const float mean = ms + pv;
    const float m2 = fmaxf(sumsq[c] - ms * sum[c], 0.f);                // sum of squared deviations from the local mean
    for (int r = 0; r < world; ++r) {
      float* dst = reinterpret_cast<float*>(slots.p[r]) + half + (size_t)rank * per;
      dst[1 + c] = mean;
      dst[1 + cap + c] = m2;
      if (c == 0) dst[0] = (float)rows;
    }
  }
  signal_wait(flags, rank, world, channel, e);
  const float* base = reinterpret_cast<const float*>(slots.p[rank]) + half;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float n = 0.f, mean = 0.f, m2 = 0.f;
    for (int r = 0; r < world; ++r) {                                   // merge (n, mean, M2) in rank order
      const float* s = base + (size_t)r * per;
      const float nb = s[0], mb = s[1 + c], m2b = s[1 + cap + c];
      const float nt = n + nb, d = mb - mean;
      mean += d * (nb / nt);
      m2 += m2b + d * d * (n * nb / nt);
      n = nt;
    }
    const float var = m2 / n;                                           // biased, over all ranks
    const float ms = mean - x[(size_t)(c / C) * group_stride + (c % C)];
    sum[c] = ms * (float)rows;
    sumsq[c] = (var + ms * ms) * (float)rows;
    if (c == 0 && total_rows_out) *total_rows_out = n;
  }
}

// Backward.  dgamma / dbeta hold this rank's sums (they stay local: they are parameter gradients and go through the gradient
// all-reduce like every other); gsum_g / gsum_b receive the sums over ALL ranks, which the input gradient needs, and
// total_rows_out the global row count.
__global__ void __launch_bounds__(1024) peer_bn_bwd_kernel(Peers flags, Peers slots, int rank, int world, unsigned* epoch, int channel,
                                                          int cap, const float* dgamma, const float* dbeta, int cols, int rows,
                                                          float* gsum_g, float* gsum_b, float* total_rows_out) {
  const size_t per = 1 + 2 * (size_t)cap;
  const unsigned e = next_epoch(epoch, channel);
  const size_t half = (size_t)(e & 1u) * MAXW * per;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float a = dgamma[c], b = dbeta[c];
    for (int r = 0; r < world; ++r) {
      float* dst = reinterpret_cast<float*>(slots.p[r]) + half + (size_t)rank * per;
      dst[1 + c] = a;
      dst[1 + cap + c] = b;
      if (c == 0) dst[0] = (float)rows;
    }
  }
  signal_wait(flags, rank, world, channel, e);
  const float* base = reinterpret_cast<const float*>(slots.p[rank]) + half;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float a = 0.f, b = 0.f, n = 0.f;
    for (int r = 0; r < world; ++r) {
      a += base[(size_t)r * per + 1 + c]; b += base[(size_t)r * per + 1 + cap + c];
      n += base[(size_t)r * per];
    }
    gsum_g[c] = a; gsum_b[c] = b;
    if (c == 0 && total_rows_out) *total_rows_out = n;
  }
}

Peers make_peers(void* const* ptrs, int world) {
  Peers P = {};
  for (int r = 0; r < world; ++r) P.p[r] = ptrs[r];
  return P;
}

}  // namespace

extern "C" {

int vdetr_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return VDETR_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  VDETR_CUDA_TRY(cudaMalloc(ptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) {          // nothing is handed out on failure
    cudaFree(*ptr);
    *ptr = nullptr;
    return (int)e;
  }
  memcpy(handle64, &h, 64);
  return 0;
}
int vdetr_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return VDETR_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  VDETR_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int vdetr_peer_close(void* ptr) {
  VDETR_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return 0;
}
int vdetr_peer_free(void* ptr) {
  VDETR_CUDA_TRY(cudaFree(ptr));
  return 0;
}
int vdetr_peer_error(int* out) {
  if (!out) return VDETR_ERR_BAD_ARG;
  VDETR_CUDA_TRY(cudaMemcpyFromSymbol(out, g_peer_error, sizeof(int)));
  return 0;
}

int vdetr_peer_barrier(void* const* flag_ptrs, int rank, int world, uint32_t* epoch, int channel, void* stream) {
  if (!flag_ptrs || !epoch || world < 1 || world > MAXW || rank < 0 || rank >= world || channel < 0 || channel >= VDETR_PEER_CHANNELS)
    return VDETR_ERR_BAD_ARG;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(make_peers(flag_ptrs, world), rank, world, epoch, channel);
  VDETR_LAUNCH_CHECK();
  return 0;
}

size_t vdetr_peer_norm_bytes(void) { return (size_t)MAXW * NORM_BLOCKS * sizeof(float); }

int vdetr_adamw_flat_peer(void* const* p_ptrs, void* const* g_ptrs, void* const* flag_ptrs, void* const* norm_ptrs, int rank, int world,
                          uint32_t* epoch, float* reduced, float* m, float* v, long long n, long long n_decay, long long lo,
                          long long hi, const float* lr, const float* step, float grad_scale_host, float max_norm, float* norm_out,
                          float beta1, float beta2, float eps, float weight_decay, void* stream) {
  if (!p_ptrs || !g_ptrs || !flag_ptrs || !norm_ptrs || !epoch || !reduced || !m || !v || !lr || !step) return VDETR_ERR_BAD_ARG;
  if (world < 1 || world > MAXW || rank < 0 || rank >= world) return VDETR_ERR_BAD_ARG;
  if (n < 0 || n_decay < 0 || n_decay > n || lo < 0 || hi < lo || hi > n || lo % 4 != 0 || (hi % 4 != 0 && hi != n)) return VDETR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PeerAdam A = {};
  A.p = make_peers(p_ptrs, world); A.g = make_peers(g_ptrs, world); A.norm_part = make_peers(norm_ptrs, world);
  A.red = reduced; A.m = m; A.v = v; A.n = n; A.n_decay = n_decay; A.lo = lo; A.hi = hi; A.lr = lr; A.step = step;
  A.norm_out = norm_out; A.gscale_host = grad_scale_host; A.max_norm = max_norm;
  A.b1 = beta1; A.b2 = beta2; A.eps = eps; A.wd = weight_decay; A.rank = rank; A.world = world;
  const Peers F = make_peers(flag_ptrs, world);
  // every rank's gradients are complete
  peer_barrier_kernel<<<1, 32, 0, st>>>(F, rank, world, epoch, 0);
  VDETR_LAUNCH_CHECK();
  peer_reduce_kernel<<<NORM_BLOCKS, 256, 0, st>>>(A);
  VDETR_LAUNCH_CHECK();
  // every rank's partial norms have arrived
  peer_barrier_kernel<<<1, 32, 0, st>>>(F, rank, world, epoch, 1);
  VDETR_LAUNCH_CHECK();
  long long blocks = ((hi - lo) / 4 + 255) / 256;
  const long long cap = (long long)vdetr_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  peer_adamw_kernel<<<(int)blocks, 256, 0, st>>>(A);
  VDETR_LAUNCH_CHECK();
  // every rank's parameters are complete: the next forward may read them, the next backward may overwrite the gradients
  peer_barrier_kernel<<<1, 32, 0, st>>>(F, rank, world, epoch, 2);
  VDETR_LAUNCH_CHECK();
  return 0;
}

size_t vdetr_peer_bn_slot_floats(int cap) { return 2 * (size_t)MAXW * (1 + 2 * (size_t)cap); }

int vdetr_peer_bn_fwd(void* const* flag_ptrs, void* const* slot_ptrs, int rank, int world, uint32_t* epoch, int cap, float* sum,
                      float* sumsq, const float* x, int cols, int groups, long long group_stride, int rows, float* total_rows_out,
                      void* stream) {
  if (!flag_ptrs || !slot_ptrs || !epoch || !sum || !sumsq || !x || cols < 1 || groups < 1 || rows < 1) return VDETR_ERR_BAD_ARG;
  if ((long long)cols * groups > cap) return VDETR_ERR_WORKSPACE;
  if (world < 1 || world > MAXW || rank < 0 || rank >= world) return VDETR_ERR_BAD_ARG;
  peer_bn_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(make_peers(flag_ptrs, world), make_peers(slot_ptrs, world), rank, world, epoch,
                                                           3, cap, sum, sumsq, x, cols, groups, group_stride, rows, total_rows_out);
  VDETR_LAUNCH_CHECK();
  return 0;
}
int vdetr_peer_bn_bwd(void* const* flag_ptrs, void* const* slot_ptrs, int rank, int world, uint32_t* epoch, int cap,
                      const float* dgamma, const float* dbeta, int cols_total, int rows, float* gsum_g, float* gsum_b,
                      float* total_rows_out, void* stream) {
  if (!flag_ptrs || !slot_ptrs || !epoch || !dgamma || !dbeta || !gsum_g || !gsum_b || cols_total < 1 || rows < 1) return VDETR_ERR_BAD_ARG;
  if (cols_total > cap) return VDETR_ERR_WORKSPACE;
  if (world < 1 || world > MAXW || rank < 0 || rank >= world) return VDETR_ERR_BAD_ARG;
  peer_bn_bwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(make_peers(flag_ptrs, world), make_peers(slot_ptrs, world), rank, world, epoch,
                                                           3, cap, dgamma, dbeta, cols_total, rows, gsum_g, gsum_b, total_rows_out);
  VDETR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
