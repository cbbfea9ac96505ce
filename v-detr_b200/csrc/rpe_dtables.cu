// dTables[i][z][y][x][h] = sum_{b,q,k} dS[b,q,k,h] * w_{i,corner}(b,q,k)        (adjoint of the bias gather;
// the reference gets it from grid_sampler_3d_backward's global atomics, vdetr_transformer.py:727-731).
//
// The scatter is 8 vertices x 8 corners x 4 heads = 256 adds per (query,key) pair into 32,000 cells of which
// a few hundred are hot, so neither global nor shared atomics per contribution are affordable (shared fp32
// atomicAdd is a CAS loop on sm_100).  Scheme:
//   * a warp owns one (scene, query, vertex) task and walks the keys 32 at a time (lane = key, Morton order);
//   * every lane writes its 32 products w_corner * dS_h to a private row of a 32x33 scratch tile;
//   * lanes are grouped by table cell ("bin" = floor of the pixel coordinate, usually 1-4 groups per step);
//     for each group lane j sums column j over the group's rows  ->  the group's 32 (corner,head) totals;
//   * the totals go into a 4-entry warp-level cache (tags are warp-uniform registers, lane j holds value j),
//     hot bins therefore stay in registers across many steps; evictions go to a per-CTA fp32 copy of the
//     tables in shared memory (CAS atomics, rare), which is added to global memory once per CTA at the end.
#include "rpe_internal.h"
#include "rpe_fast.cuh"
#include <stdlib.h>

namespace {

constexpr int DT_WARPS = 16;
constexpr int DT_THREADS = DT_WARPS * 32;
constexpr int CACHE = 8;                       // warp-level cache entries (bins held in registers)

struct DtParams {
  int B, nQ, nK, nQp, nKp, n;
  float log_scale, c1, c0;
  const float4* xyz4;       // [B][nKp]
  const float4* geo;        // [B][nQp][9]
  const float4* ds4;        // [B][nQp][nKp]  (dS of the 4 heads of one (query,key) pair)
  float* dtables;           // [8][n^3][4], zero-initialised by the caller
};

__device__ __forceinline__ void flush_slot(float* stab, int tag, float acc, int lane, int vert, int n) {
  if (tag < 0) return;
  const int n0x = (tag & 31) - 2, n0y = ((tag >> 5) & 31) - 2, n0z = ((tag >> 10) & 31) - 2;
  const int corner = lane >> 2, h = lane & 3;
  const int x = n0x + (corner & 1), y = n0y + ((corner >> 1) & 1), z = n0z + (corner >> 2);
  if ((unsigned)x < (unsigned)n && (unsigned)y < (unsigned)n && (unsigned)z < (unsigned)n && acc != 0.f)
    atomicAdd(stab + ((((size_t)vert * n + z) * n + y) * n + x) * 4 + h, acc);
}

// sum of column `lane` of the 32x33 scratch tile over the rows selected by `mask` (fully unrolled: the 32
// predicated loads are independent, so their latency overlaps)
__device__ __forceinline__ float column_sum(const float* my, int lane, unsigned mask) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int l = 0; l < 32; l += 4) {
    if (mask & (1u << l)) s0 += my[l * 33 + lane];
    if (mask & (2u << l)) s1 += my[(l + 1) * 33 + lane];
    if (mask & (4u << l)) s2 += my[(l + 2) * 33 + lane];
    if (mask & (8u << l)) s3 += my[(l + 3) * 33 + lane];
  }
  return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(DT_THREADS, 1) rpe_dtables_kernel(DtParams P) {
  extern __shared__ float dsm[];
  const int ncell4 = 8 * P.n * P.n * P.n * 4;
  float* stab = dsm;                                   // [8][n^3][4]
  float* scratch = dsm + ncell4;                       // [DT_WARPS][32][33]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < ncell4; i += DT_THREADS) stab[i] = 0.f;
  __syncthreads();
  float* my = scratch + warp * 32 * 33;
  const int sx = 16, sy = 16 * P.n, sz = 16 * P.n * P.n;

  const long tasks = (long)P.B * P.nQ * 8;
  const long gw = (long)blockIdx.x * DT_WARPS + warp, nw = (long)gridDim.x * DT_WARPS;
  for (long t = gw; t < tasks; t += nw) {
    const int vert = (int)(t & 7);
    const long bq = t >> 3;
    const int q = (int)(bq % P.nQ), b = (int)(bq / P.nQ);
    const float4* g = P.geo + ((size_t)b * P.nQp + q) * 9;
    const float* vv = reinterpret_cast<const float*>(g + 2);
    const float vx = __ldg(vv + vert * 3), vy = __ldg(vv + vert * 3 + 1), vz = __ldg(vv + vert * 3 + 2);
    const float4 rot = __ldg(g + 8);
    const float4* xrow = P.xyz4 + (size_t)b * P.nKp;
    const float4* drow = P.ds4 + ((size_t)b * P.nQp + q) * P.nKp;

    int tag[CACHE];
    float acc[CACHE];
#pragma unroll
    for (int c = 0; c < CACHE; ++c) { tag[c] = -1; acc[c] = 0.f; }
    int victim = 0;

    // software prefetch of the next step's inputs
    float4 kx_n = make_float4(0.f, 0.f, 0.f, 0.f), ds_n = kx_n;
    if (lane < P.nK) { kx_n = __ldg(xrow + lane); ds_n = __ldg(drow + lane); }
    for (int k0 = 0; k0 < P.nK; k0 += 32) {
      const int key = k0 + lane;
      const float4 kx = kx_n, ds = ds_n;
      if (key + 32 < P.nK) { kx_n = __ldg(xrow + key + 32); ds_n = __ldg(drow + key + 32); }
      int bin = -1;
      if (key < P.nK && (ds.x != 0.f || ds.y != 0.f || ds.z != 0.f || ds.w != 0.f)) {
        const float dx = vx - kx.x, dy = vy - kx.y, dz = vz - kx.z;
        const float tx = rot.x * dx - rot.y * dy, ty = rot.y * dx + rot.x * dy;
        const rpe::Axis ax = rpe::rpe_axis_fast(tx, P.log_scale, P.c1, P.c0, P.n, sx);
        const rpe::Axis ay = rpe::rpe_axis_fast(ty, P.log_scale, P.c1, P.c0, P.n, sy);
        const rpe::Axis az = rpe::rpe_axis_fast(dz, P.log_scale, P.c1, P.c0, P.n, sz);
        bin = ((az.n0 + 2) << 10) | ((ay.n0 + 2) << 5) | (ax.n0 + 2);
        float* row = my + lane * 33;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float w = ((c & 4) ? az.w1 : az.w0) * ((c & 2) ? ay.w1 : ay.w0) * ((c & 1) ? ax.w1 : ax.w0);
          row[c * 4 + 0] = w * ds.x; row[c * 4 + 1] = w * ds.y; row[c * 4 + 2] = w * ds.z; row[c * 4 + 3] = w * ds.w;
        }
      }
      __syncwarp();
      unsigned rem = __ballot_sync(0xffffffffu, bin >= 0);
      // 1. bins already cached: one masked column sum per cache entry that is hit
#pragma unroll
      for (int c = 0; c < CACHE; ++c) {
        const unsigned hit = __ballot_sync(0xffffffffu, bin == tag[c]) & rem;   // tag -1 never matches a live bin
        if (hit) { acc[c] += column_sum(my, lane, hit); rem &= ~hit; }
      }
      // 2. new bins: evict round-robin
      while (rem) {
        const int leader = __ffs(rem) - 1;
        const int bsel = __shfl_sync(0xffffffffu, bin, leader);
        const unsigned grp = __ballot_sync(0xffffffffu, bin == bsel) & rem;
        rem &= ~grp;
        const float s = column_sum(my, lane, grp);
#pragma unroll
        for (int c = 0; c < CACHE; ++c)
          if (c == victim) { flush_slot(stab, tag[c], acc[c], lane, vert, P.n); tag[c] = bsel; acc[c] = s; }
        victim = (victim + 1) & (CACHE - 1);
      }
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < CACHE; ++c) flush_slot(stab, tag[c], acc[c], lane, vert, P.n);
  }
  __syncthreads();
  for (int i = tid; i < ncell4; i += DT_THREADS) {
    const float v = stab[i];
    if (v != 0.f) atomicAdd(P.dtables + i, v);
  }
}

}  // namespace

// =====================================================================================================
// Version 2: bucket the evaluations by table cell first, then accumulate in registers.
//
// A unit is (scene b, vertex i, block of 32 queries, chunk of 1024 keys) = 32,768 evaluations.
//   phase 1  every thread computes the bin (floor of the 3 pixel coordinates) of its evaluations, stores it and
//            counts it in a shared-memory histogram (native integer atomics);
//   phase 2  exclusive scan of the histogram, then a counting-sort scatter of the evaluation ids;
//   phase 3  each warp walks a contiguous range of the sorted list, 32 evaluations per step (lane = evaluation).
//            Sorted order means a whole step normally belongs to ONE bin, so every lane simply accumulates its
//            8 corners x 4 heads into 32 private registers; only when the bin changes are the 32x32 partial sums
//            reduced across lanes (transpose through a scratch tile) and added to a per-CTA fp32 copy of table i.
// Per evaluation this costs ~4 warp-instructions instead of ~35 for the cache-based version above.
namespace v2 {

constexpr int QB = 32, KC = 1024, UNIT = QB * KC;          // evaluations per unit
constexpr int THREADS = 512, WARPS = THREADS / 32;

struct Params {
  int B, nQ, nK, nQp, nKp, n, R;     // R = n + 1 values of n0 per axis that can contribute: [-1, n-1]
  int qblocks, kchunks, units;
  float log_scale, c1, c0;
  const float4* xyz4;
  const float4* geo;
  const float4* ds4;
  float* dtables;
};

__device__ __forceinline__ int axis_n0(float d, float ls, float c1, float c0, int n) {
  float t = tc::lg2_approx(fmaf(fabsf(d), ls, 1.0f)) * c1;
  float ts = copysignf(t, d);
  ts = fminf(fmaxf(ts, -c0 - 1.5f), (float)n - c0 + 0.5f);
  float r = ((ts + c0) - 0.5f) + rpe::MAGIC;
  return __float_as_int(r) - rpe::MAGIC_BITS;
}

__global__ void __launch_bounds__(THREADS, 1) rpe_dtables_sorted_kernel(Params P) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int nbins = P.R * P.R * P.R;
  const int nbins_pad = (2 * nbins + 1 + 3) & ~3;                           // ints, keeps what follows 16-B aligned
  uint16_t* binbuf = reinterpret_cast<uint16_t*>(sm);                       // [UNIT]
  uint16_t* sorted = binbuf + UNIT;                                         // [UNIT]
  int* hist = reinterpret_cast<int*>(sorted + UNIT);                        // [nbins]  counts, then cursors
  int* offs = hist + nbins;                                                 // [nbins + 1] exclusive scan
  float* stab = reinterpret_cast<float*>(hist + nbins_pad);                 // [n^3 * 4] table of the current vertex
  float4* sgeo = reinterpret_cast<float4*>(stab + P.n * P.n * P.n * 4);     // [QB][2]: (vx,vy,vz,valid) (cos,sin,0,0)
  float* scratch = reinterpret_cast<float*>(sgeo + QB * 2);                 // [WARPS][32][33]
  __shared__ int s_total;
  __shared__ int s_next;          // phase 3: next unclaimed chunk of the sorted list

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncell4 = P.n * P.n * P.n * 4;
  float* my = scratch + warp * 32 * 33;
  const int sx = 16, sy = 16 * P.n, sz = 16 * P.n * P.n;

  // contiguous range of units per CTA, vertex-major, so that the shared table is flushed only a few times
  const int per = (P.units + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * per, u_end = min(P.units, u_begin + per);
  int cur_vert = -1;

  for (int u = u_begin; u < u_end; ++u) {
    // unit -> (vert, b, qblock, kchunk)
    int r = u;
    const int kc = r % P.kchunks; r /= P.kchunks;
    const int qb = r % P.qblocks; r /= P.qblocks;
    const int b = r % P.B;
    const int vert = r / P.B;
    const int q0 = qb * QB, k0 = kc * KC;

    if (vert != cur_vert) {
      __syncthreads();
      if (cur_vert >= 0)
        for (int i = tid; i < ncell4; i += THREADS) {
          const float v = stab[i];
          if (v != 0.f) atomicAdd(P.dtables + (size_t)cur_vert * ncell4 + i, v);
        }
      __syncthreads();
      for (int i = tid; i < ncell4; i += THREADS) stab[i] = 0.f;
      cur_vert = vert;
    }
    for (int i = tid; i < nbins; i += THREADS) hist[i] = 0;
    if (tid < QB) {
      const int q = q0 + tid;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = make_float4(1.f, 0.f, 0.f, 0.f);
      if (q < P.nQ) {
        const float4* g = P.geo + ((size_t)b * P.nQp + q) * 9;
        const float* vv = reinterpret_cast<const float*>(g + 2);
        a = make_float4(__ldg(vv + vert * 3), __ldg(vv + vert * 3 + 1), __ldg(vv + vert * 3 + 2), 1.f);
        c = __ldg(g + 8);
      }
      sgeo[tid * 2] = a; sgeo[tid * 2 + 1] = c;
    }
    __syncthreads();

    // ---- phase 1: bins + histogram.  e = ql * KC + kl ; a warp covers 32 consecutive keys of one query
    const float4* xrow = P.xyz4 + (size_t)b * P.nKp + k0;
    for (int e = tid; e < UNIT; e += THREADS) {
      const int ql = e >> 10, kl = e & (KC - 1);
      const float4 vq = sgeo[ql * 2], rot = sgeo[ql * 2 + 1];
      int bin = 0xFFFF;
      if (vq.w != 0.f && k0 + kl < P.nK) {
        const float4 kx = __ldg(xrow + kl);
        const float dx = vq.x - kx.x, dy = vq.y - kx.y, dz = vq.z - kx.z;
        const float tx = rot.x * dx - rot.y * dy, ty = rot.y * dx + rot.x * dy;
        const int nx = axis_n0(tx, P.log_scale, P.c1, P.c0, P.n) + 1;
        const int ny = axis_n0(ty, P.log_scale, P.c1, P.c0, P.n) + 1;
        const int nz = axis_n0(dz, P.log_scale, P.c1, P.c0, P.n) + 1;
        if ((unsigned)nx < (unsigned)P.R && (unsigned)ny < (unsigned)P.R && (unsigned)nz < (unsigned)P.R)
          bin = (nz * P.R + ny) * P.R + nx;
      }
      binbuf[e] = (uint16_t)bin;
      // warp-aggregated histogram update: the 32 keys of a warp mostly share a bin, and same-address shared
      // atomics serialise, so one lane per distinct bin adds the group's population
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin != 0xFFFF && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
    }
    __syncthreads();
    // ---- phase 2a: exclusive scan (warp 0: each lane owns a contiguous slice of bins), cursors = offsets
    if (warp == 0) {
      const int per_lane = (nbins + 31) / 32;
      const int lo = lane * per_lane, hi = min(nbins, lo + per_lane);
      int mine = 0;
      for (int i = lo; i < hi; ++i) mine += hist[i];
      int x = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      int run = x - mine;
      for (int i = lo; i < hi; ++i) { const int c = hist[i]; hist[i] = run; offs[i] = run; run += c; }
      if (lane == 31) { offs[nbins] = x; s_total = x; s_next = 0; }
    }
    __syncthreads();
    // ---- phase 2b: scatter
    for (int e = tid; e < UNIT; e += THREADS) {
      const int bin = binbuf[e];
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if (bin != 0xFFFF && lane == leader) base = atomicAdd(&hist[bin], __popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (bin != 0xFFFF) sorted[base + __popc(peers & ((1u << lane) - 1u))] = (uint16_t)e;
    }
    __syncthreads();
    // sorted[] no longer needs binbuf's bins except to find segment ends: re-use offs[] for that.

    // ---- phase 3: accumulate.  Warps claim chunks of CHUNK sorted entries dynamically (the cost per entry varies
    // a lot between long and short segments, so a static split leaves most warps idle at the barrier).
    constexpr int CHUNK = 512;
    const int total = s_total;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    const float4* dbase = P.ds4 + ((size_t)b * P.nQp + q0) * P.nKp + k0;
    const int corner = lane >> 2, hsel = lane & 3;

    // reduce the 32 lanes' private partial sums of bin `bin` (transpose through the scratch tile) into the table
    auto flush = [&](int bin) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { my[lane * 33 + j] = acc[j]; acc[j] = 0.f; }
      __syncwarp();
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int l = 0; l < 32; l += 4) {
        s0 += my[l * 33 + lane]; s1 += my[(l + 1) * 33 + lane];
        s2 += my[(l + 2) * 33 + lane]; s3 += my[(l + 3) * 33 + lane];
      }
      __syncwarp();
      const float tot = (s0 + s1) + (s2 + s3);
      const int bx = bin % P.R - 1, by = (bin / P.R) % P.R - 1, bz = bin / (P.R * P.R) - 1;
      const int x = bx + (corner & 1), y = by + ((corner >> 1) & 1), z = bz + (corner >> 2);
      if ((unsigned)x < (unsigned)P.n && (unsigned)y < (unsigned)P.n && (unsigned)z < (unsigned)P.n && tot != 0.f)
        atomicAdd(stab + (((z * P.n + y) * P.n + x) << 2) + hsel, tot);
    };

    for (;;) {
      int c0 = 0;
      if (lane == 0) c0 = atomicAdd(&s_next, CHUNK);
      c0 = __shfl_sync(0xffffffffu, c0, 0);
      if (c0 >= total) break;
      const int c1 = min(total, c0 + CHUNK);
      int cur_bin = -1;
      // software pipeline: the gathers of step s+1 are in flight while step s is accumulated
      int bin_n = -1;
      float4 vq_n = make_float4(0.f, 0.f, 0.f, 0.f), rot_n = vq_n, kx_n = vq_n, ds_n = vq_n;
      {
        const int p = c0 + lane;
        if (p < c1) {
          const int e = sorted[p];
          bin_n = binbuf[e];
          const int ql = e >> 10, kl = e & (KC - 1);
          vq_n = sgeo[ql * 2]; rot_n = sgeo[ql * 2 + 1];
          kx_n = __ldg(xrow + kl);
          ds_n = __ldg(dbase + (size_t)ql * P.nKp + kl);
        }
      }
      for (int p0 = c0; p0 < c1; p0 += 32) {
        const int bin = bin_n;
        const bool live = bin >= 0;
        const float4 vq = vq_n, rot = rot_n, kx = kx_n, ds = ds_n;
        bin_n = -1;
        {
          const int p = p0 + 32 + lane;
          if (p < c1) {
            const int e = sorted[p];
            bin_n = binbuf[e];
            const int ql = e >> 10, kl = e & (KC - 1);
            vq_n = sgeo[ql * 2]; rot_n = sgeo[ql * 2 + 1];
            kx_n = __ldg(xrow + kl);
            ds_n = __ldg(dbase + (size_t)ql * P.nKp + kl);
          }
        }
        // fast path: the whole step is one long segment (the common case after sorting) -> FFMA straight into
        // the lane-private accumulators, no staging of the 32 products
        {
          const unsigned livemask = __ballot_sync(0xffffffffu, live);
          const int lead = __ffs(livemask) - 1;
          const int b0 = __shfl_sync(0xffffffffu, bin, lead < 0 ? 0 : lead);
          const bool uniform = livemask != 0u && __all_sync(0xffffffffu, !live || bin == b0);
          if (uniform && (b0 == cur_bin || __popc(livemask) >= 12)) {
            if (b0 != cur_bin) {
              if (cur_bin >= 0) flush(cur_bin);
              cur_bin = b0;
            }
            if (live) {
              const float dx = vq.x - kx.x, dy = vq.y - kx.y, dz = vq.z - kx.z;
              const float tx = rot.x * dx - rot.y * dy, ty = rot.y * dx + rot.x * dy;
              const rpe::Axis ax = rpe::rpe_axis_fast(tx, P.log_scale, P.c1, P.c0, P.n, sx);
              const rpe::Axis ay = rpe::rpe_axis_fast(ty, P.log_scale, P.c1, P.c0, P.n, sy);
              const rpe::Axis az = rpe::rpe_axis_fast(dz, P.log_scale, P.c1, P.c0, P.n, sz);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float w = ((k & 4) ? az.w1 : az.w0) * ((k & 2) ? ay.w1 : ay.w0) * ((k & 1) ? ax.w1 : ax.w0);
                acc[k * 4 + 0] = fmaf(w, ds.x, acc[k * 4 + 0]); acc[k * 4 + 1] = fmaf(w, ds.y, acc[k * 4 + 1]);
                acc[k * 4 + 2] = fmaf(w, ds.z, acc[k * 4 + 2]); acc[k * 4 + 3] = fmaf(w, ds.w, acc[k * 4 + 3]);
              }
            }
            continue;
          }
        }
        float c[32];
        if (live) {
          const float dx = vq.x - kx.x, dy = vq.y - kx.y, dz = vq.z - kx.z;
          const float tx = rot.x * dx - rot.y * dy, ty = rot.y * dx + rot.x * dy;
          const rpe::Axis ax = rpe::rpe_axis_fast(tx, P.log_scale, P.c1, P.c0, P.n, sx);
          const rpe::Axis ay = rpe::rpe_axis_fast(ty, P.log_scale, P.c1, P.c0, P.n, sy);
          const rpe::Axis az = rpe::rpe_axis_fast(dz, P.log_scale, P.c1, P.c0, P.n, sz);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float w = ((k & 4) ? az.w1 : az.w0) * ((k & 2) ? ay.w1 : ay.w0) * ((k & 1) ? ax.w1 : ax.w0);
            c[k * 4 + 0] = w * ds.x; c[k * 4 + 1] = w * ds.y; c[k * 4 + 2] = w * ds.z; c[k * 4 + 3] = w * ds.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) c[k] = 0.f;
        }
        // segments of this step, in sorted (= lane) order.  Long segments (>= 12 lanes, or the continuation of the
        // bin being accumulated) go to the lane-private registers; the long tail of nearly empty cells is summed
        // through the scratch tile (rows of one segment are contiguous) and added straight to the table.
        unsigned todo = __ballot_sync(0xffffffffu, live), longmask = 0u;
        bool staged = false;
        while (todo) {                                   // pass A: short segments
          const int leader = __ffs(todo) - 1;
          const int bsel = __shfl_sync(0xffffffffu, bin, leader);
          const unsigned grp = __ballot_sync(0xffffffffu, live && bin == bsel);
          todo &= ~grp;
          const int cnt = __popc(grp);
          if (bsel == cur_bin || cnt >= 12) { longmask |= grp; continue; }
          if (!staged) {
#pragma unroll
            for (int j = 0; j < 32; ++j) my[lane * 33 + j] = c[j];
            __syncwarp();
            staged = true;
          }
          float ssum = 0.f;
          for (int l = leader; l < leader + cnt; ++l) ssum += my[l * 33 + lane];
          const int bx = bsel % P.R - 1, by = (bsel / P.R) % P.R - 1, bz = bsel / (P.R * P.R) - 1;
          const int x = bx + (corner & 1), y = by + ((corner >> 1) & 1), z = bz + (corner >> 2);
          if ((unsigned)x < (unsigned)P.n && (unsigned)y < (unsigned)P.n && (unsigned)z < (unsigned)P.n && ssum != 0.f)
            atomicAdd(stab + (((z * P.n + y) * P.n + x) << 2) + hsel, ssum);
        }
        if (staged) __syncwarp();
        todo = longmask;
        while (todo) {                                   // pass B: long segments
          const int leader = __ffs(todo) - 1;
          const int bsel = __shfl_sync(0xffffffffu, bin, leader);
          const unsigned grp = __ballot_sync(0xffffffffu, live && bin == bsel) & longmask;
          todo &= ~grp;
          if (bsel != cur_bin) {
            if (cur_bin >= 0) flush(cur_bin);
            cur_bin = bsel;
          }
          if ((grp >> lane) & 1u) {
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] += c[k];
          }
        }
      }
      if (cur_bin >= 0) flush(cur_bin);
    }
    __syncthreads();     // before the next unit overwrites hist / binbuf / sorted / sgeo
  }
  __syncthreads();
  if (cur_vert >= 0)
    for (int i = tid; i < ncell4; i += THREADS) {
      const float v = stab[i];
      if (v != 0.f) atomicAdd(P.dtables + (size_t)cur_vert * ncell4 + i, v);
    }
}

size_t smem_bytes(int n) {
  const int R = n + 1, nbins = R * R * R;
  size_t o = (size_t)UNIT * 2 * 2;
  o += (size_t)((2 * nbins + 1 + 3) & ~3) * 4;
  o += (size_t)n * n * n * 16 + QB * 2 * 16 + (size_t)WARPS * 32 * 33 * 4;
  return o;
}

}  // namespace v2

// ds4 [B][nQp][nKp] float4, xyz4 / geo as produced by vdetr_pack_kernel.  dtables is zeroed here.
int rpe_dtables_launch(const VdetrXattnShape* s, int nQp, int nKp, const float4* xyz4, const float4* geo, const float4* ds4,
                       float* dtables, cudaStream_t st) {
  const int n = s->grid_n;
  const size_t tbytes = (size_t)8 * n * n * n * 4 * sizeof(float);
  VDETR_CUDA_TRY(cudaMemsetAsync(dtables, 0, tbytes, st));
  if (s->B == 0 || s->nQ == 0 || s->nK == 0) return 0;
  DtParams P = {};
  P.B = s->B; P.nQ = s->nQ; P.nK = s->nK; P.nQp = nQp; P.nKp = nKp; P.n = n;
  P.log_scale = s->log_scale;
  P.c1 = (float)n / (2.0f * 3.0f * s->max_value);
  P.c0 = 0.5f * (float)(n - 1);
  P.xyz4 = xyz4; P.geo = geo; P.ds4 = ds4; P.dtables = dtables;
  const char* force = getenv("VDETR_DT_IMPL");
  const size_t smem2 = v2::smem_bytes(n);
  if (smem2 <= 232448 && (n + 1) * (n + 1) * (n + 1) < 0xFFFF && !(force && force[0] == '1')) {
    v2::Params Q = {};
    Q.B = s->B; Q.nQ = s->nQ; Q.nK = s->nK; Q.nQp = nQp; Q.nKp = nKp; Q.n = n; Q.R = n + 1;
    Q.qblocks = (s->nQ + v2::QB - 1) / v2::QB;
    Q.kchunks = (s->nK + v2::KC - 1) / v2::KC;
    Q.units = 8 * s->B * Q.qblocks * Q.kchunks;
    Q.log_scale = P.log_scale; Q.c1 = P.c1; Q.c0 = P.c0;
    Q.xyz4 = xyz4; Q.geo = geo; Q.ds4 = ds4; Q.dtables = dtables;
    VDETR_CUDA_TRY(cudaFuncSetAttribute(v2::rpe_dtables_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    const int grid2 = Q.units < vdetr_num_sms() ? Q.units : vdetr_num_sms();
    {
      VdetrTimingScope timing(VDETR_T_DTABLES, st);
      v2::rpe_dtables_sorted_kernel<<<grid2, v2::THREADS, smem2, st>>>(Q);
    }
    VDETR_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = tbytes + (size_t)DT_WARPS * 32 * 33 * sizeof(float);
  if (smem > 232448) return VDETR_ERR_UNSUPPORTED;
  VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_dtables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long tasks = (long)s->B * s->nQ * 8;
  long want = (tasks + DT_WARPS - 1) / DT_WARPS;
  int grid = (int)(want < vdetr_num_sms() ? want : vdetr_num_sms());
  {
    VdetrTimingScope timing(VDETR_T_DTABLES, st);
    rpe_dtables_kernel<<<grid, DT_THREADS, smem, st>>>(P);
  }
  VDETR_LAUNCH_CHECK();
  return 0;
}

// C-ABI helper behind vdetr_rpe_dtables: dense dS [B,nQ,nK,4] -> dTables (packs xyz / geometry itself).
size_t rpe_dtables_workspace(const VdetrXattnShape* s) {
  return vdetr_align_up((size_t)s->B * s->nK * 16, 1024) + vdetr_align_up((size_t)s->B * s->nQ * 9 * 16, 1024);
}
int rpe_dtables_dense(const VdetrXattnShape* s, const float* xyz, const float* ref, const float* ang, const float* ds4,
                      float* dtables, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < rpe_dtables_workspace(s)) return VDETR_ERR_WORKSPACE;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  VdetrPack pk = {};
  pk.B = s->B; pk.nQ = s->nQ; pk.nK = s->nK; pk.nQp = s->nQ; pk.nKp = s->nK; pk.kvh = 1; pk.has_bias = 1;
  pk.xyz = xyz; pk.ref = ref; pk.ang = s->rotate ? ang : nullptr;
  pk.xyz4 = reinterpret_cast<float4*>(w);
  pk.geo = reinterpret_cast<float4*>(w + vdetr_align_up((size_t)s->B * s->nK * 16, 1024));
  vdetr_pack_kernel<<<vdetr_num_sms(), 256, 0, st>>>(pk);
  VDETR_LAUNCH_CHECK();
  return rpe_dtables_launch(s, s->nQ, s->nK, pk.xyz4, pk.geo, reinterpret_cast<const float4*>(ds4), dtables, st);
}
