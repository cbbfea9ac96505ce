// pointnet2 seed ops for sm_100a: furthest point sampling, gather, ball query, grouping.
//
// Replaces /root/reference/third_party/pointnet2/_ext_src/src/{sampling,ball_query,group_points}_gpu.cu.
// Index outputs are bit-exact with the reference kernels (same FP32 operation order, same tie-breaks),
// but the execution model is different:
//
//  * FPS: the reference runs one 512-thread block per scene and re-reads xyz + temp from global memory
//    in each of the M-1 strictly sequential rounds (sampling_gpu.cu:92-113) followed by a 9-level
//    shared-memory tree (:118-171).  Here a thread-block CLUSTER owns a scene: every point and its running
//    min-distance live in registers for the whole kernel (N*16 B spread over up to 16 SMs), a round is
//    FSUB/FMUL/FFMA/FMNMX on registers + one warp-shuffle arg-max + one DSMEM exchange, and nothing but
//    the winning index touches global memory.
//
//    Exactness.  The reference winner of a round is arg-max of temp[k] with ties broken by
//      (a) inside reference-thread t (k = t mod S): smallest k          (strict '>' at sampling_gpu.cu:111)
//      (b) between reference-threads: the tree keeps idx1 on ties (:62-68), i.e. the thread whose
//          log2(S)-bit id is smallest after BIT REVERSAL (level with stride 1 decides first).
//    Both rules together are the total order  slot(k) = bitrev(k mod S) * ceil(N/S) + k div S.
//    We hand out points to threads in increasing slot order, so "first strict max" in a thread, lower lane
//    in a shuffle, lower warp and lower CTA rank all agree with the reference's choice.
//
//  * ball_query: the reference lets each thread walk all N points in global memory (ball_query_gpu.cu:29).
//    Here the CTA streams xyz through shared memory in tiles (broadcast LDS.128) with block-wide early exit.
#include <cooperative_groups.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int FPS_THREADS = 512;
constexpr int FPS_WARPS = FPS_THREADS / 32;

struct FpsGeom {
  int n, m;
  int S;      // opt_n_threads(n) of the reference (include/cuda_utils.h:17-21)
  int L;      // log2(S)
  int cnt;    // ceil(n / S)
  int slots;  // S * cnt
};

__host__ __device__ inline uint32_t bitrev_l(uint32_t v, int L) {
  uint32_t r = 0;
  for (int i = 0; i < L; ++i) r |= ((v >> i) & 1u) << (L - 1 - i);
  return r;
}
__device__ __forceinline__ int slot_to_k(uint32_t slot, const FpsGeom& g) {
  uint32_t r = slot / (uint32_t)g.cnt, q = slot - r * (uint32_t)g.cnt;
  uint32_t br = g.L ? (__brev(r) >> (32 - g.L)) : 0u;
  return (int)(br + (q << g.L));
}

struct Cand {
  unsigned long long key;  // (bits(dist) << 32) | ~slot ; 0 = no candidate
  float x, y, z;
  float pad;
};

__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long k, int width = 32) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    if (o < width) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
      k = other > k ? other : k;
    }
  }
  return k;
}
__device__ __forceinline__ void cand_max(Cand& a, unsigned long long k, float x, float y, float z) {
  if (k > a.key) { a.key = k; a.x = x; a.y = y; a.z = z; }
}
__device__ __forceinline__ void warp_argmax(Cand& c, int width = 32) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    if (o < width) {
      unsigned long long k = __shfl_xor_sync(0xffffffffu, c.key, o);
      float x = __shfl_xor_sync(0xffffffffu, c.x, o);
      float y = __shfl_xor_sync(0xffffffffu, c.y, o);
      float z = __shfl_xor_sync(0xffffffffu, c.z, o);
      cand_max(c, k, x, y, z);
    }
  }
}

// One cluster (CLUSTER CTAs) per scene, PPT points per thread.  Coordinates + running min-distance live in
// registers; a float4 copy of the CTA's coordinates sits in shared memory so that only the 64-bit key travels
// through the shuffles and the winner's xyz is fetched once per round by one thread.
// geometry of a scene of n points, computed on the device for the ragged variant: floor(log2 n) equals the host's
// (int)(log(n) / log(2)) for every n (checked exhaustively below 2^22; S is clamped to 512 anyway)
__device__ __forceinline__ FpsGeom device_geom(int n, int m) {
  FpsGeom g;
  g.n = n; g.m = m;
  int L = n > 0 ? 31 - __clz(n) : 0;
  if (L > 9) L = 9;
  g.L = L; g.S = 1 << L;
  g.cnt = (n + g.S - 1) / g.S;
  g.slots = g.S * g.cnt;
  return g;
}

// offsets == nullptr: B scenes of g.n points each, xyz [B][n][3].
// offsets != nullptr: ragged batch, scene b = rows [offsets[b], offsets[b+1]) of xyz [total][3]; every scene is sampled
// exactly like a B = 1 call on its own points (the reference's per-scene loop, models/model_vdetr.py:282-316).
template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offsets, int32_t* __restrict__ idxs, FpsGeom g,
                   int cluster_size) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int scene = blockIdx.x / cluster_size;
  const float* pts = xyz + (size_t)scene * g.n * 3;
  int32_t* out = idxs + (size_t)scene * g.m;
  if (offsets) {
    const int o0 = __ldg(offsets + scene), n = __ldg(offsets + scene + 1) - o0;
    pts = xyz + (size_t)o0 * 3;
    const int m = g.m;
    g = device_geom(n, m);
    if (n <= 0 || (long long)g.slots > (long long)cluster_size * FPS_THREADS * PPT) {
      // empty scene, or more points than the caller's bound: mark the row invalid instead of sampling garbage
      // (cluster-uniform branch: no cluster.sync() has been executed yet)
      if (rank == 0)
        for (int j = threadIdx.x; j < m; j += FPS_THREADS) out[j] = n <= 0 ? 0 : -1;
      return;
    }
  }

  extern __shared__ float4 spts[];                 // [FPS_THREADS * PPT] this CTA's points, slot order
  __shared__ unsigned long long warp_key[2][FPS_WARPS];
  __shared__ Cand cta_cand[2][16];                 // written by every CTA of the cluster through DSMEM

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cta_slot0 = (uint32_t)rank * FPS_THREADS * PPT;
  const uint32_t slot0 = cta_slot0 + (uint32_t)tid * PPT;

  float px[PPT], py[PPT], pz[PPT], pt[PPT];
  uint32_t valid = 0;
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    uint32_t s = slot0 + p;
    float x = 0.f, y = 0.f, z = 0.f;
    pt[p] = 1e10f;                                            // sampling.cpp:75-77
    if (s < (uint32_t)g.slots) {
      int k = slot_to_k(s, g);
      if (k < g.n) {
        x = pts[k * 3 + 0]; y = pts[k * 3 + 1]; z = pts[k * 3 + 2];
        float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));   // nvcc contracts as FMUL y,y; FFMA x,x; FFMA z,z (:103)
        if (!((double)mag <= 1e-3)) valid |= 1u << p;                    // :104 (double literal)
      }
    }
    px[p] = x; py[p] = y; pz[p] = z;
    spts[tid * PPT + p] = make_float4(x, y, z, 0.f);
  }
  float ox = pts[0], oy = pts[1], oz = pts[2];                // old = 0 (:88-96)
  if (rank == 0 && tid == 0) out[0] = 0;
  if (cluster_size > 1) cluster.sync();   // peers must be resident before their smem is written
  else __syncthreads();

  for (int j = 1; j < g.m; ++j) {
    float best = -1.f;
    int bp = 0;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      if (valid & (1u << p)) {
        float dx = __fsub_rn(px[p], ox), dy = __fsub_rn(py[p], oy), dz = __fsub_rn(pz[p], oz);
        float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));   // :107-108
        float d2 = fminf(d, pt[p]);                                           // :110
        pt[p] = d2;
        if (d2 > best) { best = d2; bp = p; }                                 // :111-112
      }
    }
    unsigned long long key = (best < 0.f) ? 0ull
                         : (((unsigned long long)__float_as_uint(best) << 32) |
                            (unsigned long long)(0xFFFFFFFFu - (slot0 + (uint32_t)bp)));
    key = warp_max_key(key);
    const int par = j & 1;
    if (lane == 0) warp_key[par][warp] = key;
    __syncthreads();
    key = warp_max_key(warp_key[par][lane & (FPS_WARPS - 1)], FPS_WARPS);
    Cand b;
    b.key = key; b.x = b.y = b.z = b.pad = 0.f;
    if (cluster_size == 1) {
      if (key != 0ull) {
        float4 w = spts[(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull)) - cta_slot0];
        b.x = w.x; b.y = w.y; b.z = w.z;
      }
    } else {
      if (warp == 0 && lane < cluster_size) {
        if (key != 0ull) {
          float4 w = spts[(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull)) - cta_slot0];
          b.x = w.x; b.y = w.y; b.z = w.z;
        }
        Cand* remote = cluster.map_shared_rank(&cta_cand[par][rank], lane);
        *remote = b;
      }
      cluster.sync();
      Cand e;
      e.key = 0ull; e.x = e.y = e.z = e.pad = 0.f;
      if ((lane & 15) < cluster_size) e = cta_cand[par][lane & 15];
      warp_argmax(e, 16);
      b = e;
    }
    if (b.key == 0ull) {            // no candidate at all: old = 0 (besti = 0 in every thread)
      ox = pts[0]; oy = pts[1]; oz = pts[2];
      if (rank == 0 && tid == 0) out[j] = 0;
    } else {
      ox = b.x; oy = b.y; oz = b.z;
      if (rank == 0 && tid == 0) out[j] = slot_to_k(0xFFFFFFFFu - (uint32_t)(b.key & 0xFFFFFFFFull), g);
    }
  }
  if (cluster_size > 1) cluster.sync();   // no CTA may exit while peers can still write its smem
}

// Fallback for very large N: temp lives in global memory (workspace), one 1024-thread CTA per scene.
__global__ void __launch_bounds__(1024, 1)
fps_generic_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offsets, float* __restrict__ temp,
                   int32_t* __restrict__ idxs, FpsGeom g) {
  const int scene = blockIdx.x;
  const float* pts = xyz + (size_t)scene * g.n * 3;
  float* tmp = temp + (size_t)scene * g.n;
  int32_t* out = idxs + (size_t)scene * g.m;
  if (offsets) {                       // ragged batch: see fps_cluster_kernel
    const int o0 = __ldg(offsets + scene), n = __ldg(offsets + scene + 1) - o0;
    pts = xyz + (size_t)o0 * 3;
    tmp = temp + o0;
    const int m = g.m;
    g = device_geom(n, m);
    if (n <= 0) {
      for (int j = threadIdx.x; j < m; j += 1024) out[j] = 0;
      return;
    }
  }
  __shared__ Cand warp_cand[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < g.n; k += 1024) tmp[k] = 1e10f;
  float ox = pts[0], oy = pts[1], oz = pts[2];
  if (tid == 0) out[0] = 0;
  __syncthreads();
  for (int j = 1; j < g.m; ++j) {
    Cand c; c.key = 0ull; c.x = c.y = c.z = c.pad = 0.f;
    for (int k = tid; k < g.n; k += 1024) {
      float x = pts[k * 3], y = pts[k * 3 + 1], z = pts[k * 3 + 2];
      float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
      if ((double)mag <= 1e-3) continue;
      float dx = __fsub_rn(x, ox), dy = __fsub_rn(y, oy), dz = __fsub_rn(z, oz);
      float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      float d2 = fminf(d, tmp[k]);
      tmp[k] = d2;
      if (d2 > -1.f) {
        uint32_t slot = bitrev_l((uint32_t)k & (uint32_t)(g.S - 1), g.L) * (uint32_t)g.cnt + ((uint32_t)k >> g.L);
        unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(0xFFFFFFFFu - slot);
        cand_max(c, key, x, y, z);
      }
    }
    warp_argmax(c);
    const int par = j & 1;
    if (lane == 0) warp_cand[par][warp] = c;
    __syncthreads();
    Cand b = warp_cand[par][lane];
    warp_argmax(b);
    if (b.key == 0ull) { ox = pts[0]; oy = pts[1]; oz = pts[2]; if (tid == 0) out[j] = 0; }
    else { ox = b.x; oy = b.y; oz = b.z; if (tid == 0) out[j] = slot_to_k(0xFFFFFFFFu - (uint32_t)(b.key & 0xFFFFFFFFull), g); }
  }
}

FpsGeom make_geom(int n, int m) {
  FpsGeom g;
  g.n = n; g.m = m;
  const int pow_2 = (int)(log((double)n) / log(2.0));          // cuda_utils.h:17-21, same expression
  int S = 1 << pow_2;
  if (S > 512) S = 512;
  if (S < 1) S = 1;
  g.S = S;
  int L = 0;
  while ((1 << L) < S) ++L;
  g.L = L;
  g.cnt = (n + S - 1) / S;
  g.slots = g.S * g.cnt;
  return g;
}

struct FpsPlan { int cluster; int ppt; };   // cluster == 0 -> generic kernel
// Smallest cluster that keeps <= 8 points per thread (a round's issue time grows with PPT); clusters of 16
// are non-portable, so they are only planned when `allow16` (and the caller retries without on failure).
FpsPlan fps_plan(const FpsGeom& g, bool allow16) {
  const int ppts[4] = {4, 8, 16, 24};
  const int cmax = allow16 ? 16 : 8;
  for (int c = 1; c <= cmax; c *= 2)
    for (int i = 0; i < 2; ++i)
      if ((long long)c * FPS_THREADS * ppts[i] >= g.slots) return FpsPlan{c, ppts[i]};
  for (int i = 2; i < 4; ++i)
    if ((long long)cmax * FPS_THREADS * ppts[i] >= g.slots) return FpsPlan{cmax, ppts[i]};
  return FpsPlan{0, 0};
}

// max_active != nullptr: do not launch, report how many clusters of this shape can be resident at once
template <int PPT>
int launch_fps_cluster(const float* xyz, const int32_t* offsets, int32_t* idx, const FpsGeom& g, int B, int cluster, cudaStream_t st,
                       int* max_active = nullptr) {
  auto kern = fps_cluster_kernel<PPT>;
  if (cluster > 8)
    VDETR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cluster));
  cfg.blockDim = dim3(FPS_THREADS);
  const size_t smem = (size_t)FPS_THREADS * PPT * sizeof(float4);
  VDETR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (max_active) {
    VDETR_CUDA_TRY(cudaOccupancyMaxActiveClusters(max_active, kern, &cfg));
    return 0;
  }
  VDETR_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, xyz, offsets, idx, g, cluster));
  return 0;
}
int launch_fps_plan(const FpsPlan& p, const float* xyz, const int32_t* offsets, int32_t* idx, const FpsGeom& g, int B,
                    cudaStream_t st, int* max_active = nullptr) {
  switch (p.ppt) {
    case 4: return launch_fps_cluster<4>(xyz, offsets, idx, g, B, p.cluster, st, max_active);
    case 8: return launch_fps_cluster<8>(xyz, offsets, idx, g, B, p.cluster, st, max_active);
    case 16: return launch_fps_cluster<16>(xyz, offsets, idx, g, B, p.cluster, st, max_active);
    default: return launch_fps_cluster<24>(xyz, offsets, idx, g, B, p.cluster, st, max_active);
  }
}

// ------------------------------------------------------------------------------------------ gather / group
__global__ void gather_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int C, int N, int M,
                              float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float* src = points + ((size_t)b * C + c) * N;
  float* dst = out + ((size_t)b * C + c) * M;
  const int32_t* ib = idx + (size_t)b * M;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) dst[j] = src[ib[j]];
}
__global__ void gather_grad_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int C, int N, int M,
                                   float* __restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float* src = grad_out + ((size_t)b * C + c) * M;
  float* dst = grad_points + ((size_t)b * C + c) * N;
  const int32_t* ib = idx + (size_t)b * M;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) atomicAdd(dst + ib[j], src[j]);
}
// out[b,c,j,s] = points[b,c,idx[b,j,s]] ; one thread per output element, (j,s) fastest
__global__ void group_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int C, int N, int MS,
                             float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float* src = points + ((size_t)b * C + c) * N;
  float* dst = out + ((size_t)b * C + c) * MS;
  const int32_t* ib = idx + (size_t)b * MS;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < MS; j += gridDim.x * blockDim.x) dst[j] = src[ib[j]];
}
__global__ void group_grad_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int C, int N, int MS,
                                  float* __restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float* src = grad_out + ((size_t)b * C + c) * MS;
  float* dst = grad_points + ((size_t)b * C + c) * N;
  const int32_t* ib = idx + (size_t)b * MS;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < MS; j += gridDim.x * blockDim.x) atomicAdd(dst + ib[j], src[j]);
}

// ------------------------------------------------------------------------------------------ ball query
// A warp owns a centre and tests 32 consecutive points per step; hits are appended in point order with a ballot /
// prefix-popcount, which reproduces the reference's sequential scan (first `nsample` hits in index order, all slots
// pre-filled with the first hit: ball_query_gpu.cu:25-41) exactly.  The CTA's 8 centres share 1024-point tiles of xyz
// in shared memory; the scan stops as soon as every centre of the CTA is full.
constexpr int BQ_WARPS = 8;
constexpr int BQ_THREADS = BQ_WARPS * 32;
constexpr int BQ_TILE = 1024;
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz, int N, int M, float radius, int nsample,
                  int32_t* __restrict__ idx) {
  __shared__ float4 tile[BQ_TILE];
  const int b = blockIdx.y;
  const float* pts = xyz + (size_t)b * N * 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * BQ_WARPS + warp;
  const bool active = j < M;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (active) {
    const float* c = new_xyz + ((size_t)b * M + j) * 3;
    cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
  }
  int32_t* row = idx + ((size_t)b * M + (active ? j : 0)) * nsample;
  const float radius2 = __fmul_rn(radius, radius);                 // ball_query_gpu.cu:25
  int cnt = active ? 0 : nsample;                                  // warp-uniform
  if (active)                                                      // rows without a hit stay zero: ball_query.cpp:22-24
    for (int l = lane; l < nsample; l += 32) row[l] = 0;
  __syncwarp();
  for (int base = 0; base < N; base += BQ_TILE) {
    const int len = min(BQ_TILE, N - base);
    __syncthreads();
    for (int t = threadIdx.x; t < len; t += BQ_THREADS) {
      const float* p = pts + (size_t)(base + t) * 3;
      tile[t] = make_float4(p[0], p[1], p[2], 0.f);
    }
    __syncthreads();
    for (int t0 = 0; t0 < len && cnt < nsample; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      if (t < len) {
        const float4 p = tile[t];
        const float dx = __fsub_rn(cx, p.x), dy = __fsub_rn(cy, p.y), dz = __fsub_rn(cz, p.z);
        const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));     // :33-35, FMA-contracted
        hit = d2 < radius2;                                                        // strict, :36
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (mask) {
        if (cnt == 0) {                                                            // :37-41: first hit fills every slot
          const int first = base + t0 + __ffs(mask) - 1;
          for (int l = lane; l < nsample; l += 32) row[l] = first;
          __syncwarp();
        }
        const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < nsample) row[pos] = base + t;
        cnt = min(nsample, cnt + __popc(mask));
      }
    }
    if (__syncthreads_and(cnt >= nsample)) break;
  }
}

}  // namespace

// ============================================================================================= C ABI
extern "C" {

size_t vdetr_pn2_fps_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  FpsGeom g = make_geom(N, M);
  FpsPlan p = fps_plan(g, false);
  return p.cluster == 0 ? (size_t)B * N * sizeof(float) : 0;
}

// Plans and launches FPS for B scenes whose geometry is bounded by g (exact for the dense call, the caller's max_n for
// the ragged one).  temp_floats = total number of points (workspace of the generic kernel).
static int fps_dispatch(const float* xyz, const int32_t* offsets, int B, const FpsGeom& g, size_t temp_floats, int32_t* idx,
                        void* workspace, size_t workspace_bytes, cudaStream_t st) {
  for (int attempt = 0; attempt < 2; ++attempt) {
    // 16-CTA clusters (non-portable) need 16 free SMs in one GPC: they are used only when all B clusters can be resident
    // at once (occupancy query below) -- otherwise half the scenes wait for a second wave and B = 8 costs 2 x B = 1.
    const bool allow16 = (attempt == 0) && (B * 16 <= vdetr_num_sms());
    FpsPlan p = fps_plan(g, allow16);
    if (const char* dbg = getenv("VDETR_FPS_PLAN")) {          // debug override "cluster,ppt" (tests only)
      int c = 0, pp = 0;
      if (sscanf(dbg, "%d,%d", &c, &pp) == 2 && (long long)c * FPS_THREADS * pp >= g.slots) p = FpsPlan{c, pp};
    }
    if (p.cluster == 0) {
      if (!workspace || workspace_bytes < temp_floats * sizeof(float)) return VDETR_ERR_WORKSPACE;
      fps_generic_kernel<<<B, 1024, 0, st>>>(xyz, offsets, (float*)workspace, idx, g);
      VDETR_LAUNCH_CHECK();
      return 0;
    }
    if (p.cluster > 8) {
      int active = 0;
      const int qrc = launch_fps_plan(p, xyz, offsets, idx, g, B, st, &active);
      if (qrc != 0 || active < B) {
        (void)cudaGetLastError();
        continue;                                              // portable plan: 8-CTA clusters, more points per thread
      }
    }
    const int rc = launch_fps_plan(p, xyz, offsets, idx, g, B, st);
    if (rc == 0) ++g_vdetr_launches;
    if (rc == 0 || p.cluster <= 8) return rc;
    (void)cudaGetLastError();   // non-portable cluster refused: retry with the portable plan
  }
  return VDETR_ERR_UNSUPPORTED;
}

int vdetr_pn2_fps(const float* xyz, int B, int N, int M, int32_t* idx, void* workspace, size_t workspace_bytes,
                  void* stream) {
  if (B < 0 || N < 0 || M < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || M == 0) return 0;
  if (N == 0 || !xyz || !idx) return VDETR_ERR_BAD_ARG;
  return fps_dispatch(xyz, nullptr, B, make_geom(N, M), (size_t)B * N, idx, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t vdetr_pn2_fps_ragged_workspace_bytes(int B, int total_n, int max_n, int M) {
  if (B <= 0 || total_n <= 0 || max_n <= 0 || M <= 0) return 0;
  FpsGeom g = make_geom(max_n, M);
  FpsPlan p = fps_plan(g, false);
  return p.cluster == 0 ? (size_t)total_n * sizeof(float) : 0;
}

int vdetr_pn2_fps_ragged(const float* xyz, const int32_t* offsets, int B, int total_n, int max_n, int M, int32_t* idx,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || total_n < 0 || max_n < 0 || M < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || M == 0) return 0;
  if (!xyz || !offsets || !idx || max_n == 0) return VDETR_ERR_BAD_ARG;
  return fps_dispatch(xyz, offsets, B, make_geom(max_n, M), (size_t)total_n, idx, workspace, workspace_bytes, (cudaStream_t)stream);
}

static inline dim3 rows_grid(int work, int C, int B) {
  int gx = (work + 255) / 256;
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)C, (unsigned)B);
}

int vdetr_pn2_gather(const float* points, const int32_t* idx, int B, int C, int N, int M, float* out, void* stream) {
  if (B < 0 || C < 0 || N < 0 || M < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || C == 0 || M == 0) return 0;
  if (C > 65535 || B > 65535) return VDETR_ERR_UNSUPPORTED;
  gather_kernel<<<rows_grid(M, C, B), 256, 0, (cudaStream_t)stream>>>(points, idx, C, N, M, out);
  VDETR_LAUNCH_CHECK();
  return 0;
}
int vdetr_pn2_gather_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int M, float* grad_points,
                          void* stream) {
  if (B < 0 || C < 0 || N < 0 || M < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || C == 0 || M == 0) return 0;
  if (C > 65535 || B > 65535) return VDETR_ERR_UNSUPPORTED;
  gather_grad_kernel<<<rows_grid(M, C, B), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, M, grad_points);
  VDETR_LAUNCH_CHECK();
  return 0;
}
int vdetr_pn2_group(const float* points, const int32_t* idx, int B, int C, int N, int M, int S, float* out, void* stream) {
  if (B < 0 || C < 0 || N < 0 || M < 0 || S < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || C == 0 || M == 0 || S == 0) return 0;
  if (C > 65535 || B > 65535) return VDETR_ERR_UNSUPPORTED;
  group_kernel<<<rows_grid(M * S, C, B), 256, 0, (cudaStream_t)stream>>>(points, idx, C, N, M * S, out);
  VDETR_LAUNCH_CHECK();
  return 0;
}
int vdetr_pn2_group_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int M, int S, float* grad_points,
                         void* stream) {
  if (B < 0 || C < 0 || N < 0 || M < 0 || S < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || C == 0 || M == 0 || S == 0) return 0;
  if (C > 65535 || B > 65535) return VDETR_ERR_UNSUPPORTED;
  group_grad_kernel<<<rows_grid(M * S, C, B), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, M * S, grad_points);
  VDETR_LAUNCH_CHECK();
  return 0;
}

int vdetr_pn2_ball_query(const float* new_xyz, const float* xyz, int B, int N, int M, float radius, int nsample,
                         int32_t* idx, void* stream) {
  if (B < 0 || N < 0 || M < 0 || nsample < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || M == 0 || nsample == 0) return 0;
  if (B > 65535) return VDETR_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((M + BQ_WARPS - 1) / BQ_WARPS), (unsigned)B);
  ball_query_kernel<<<grid, BQ_THREADS, 0, (cudaStream_t)stream>>>(new_xyz, xyz, N, M, radius, nsample, idx);
  VDETR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
