// dTables[i][z][y][x][h] = sum_{b,q,k} dS[b,q,k,h] * w_{i,corner}(b,q,k)        (adjoint of the bias gather;
// the reference gets it from grid_sampler_3d_backward's global atomics, vdetr_transformer.py:727-731).
//
// The scatter is 8 vertices x 8 corners x 4 heads = 256 adds per (query,key) pair into 8 x 1000 cells of which a
// few hundred are hot.  Shared-memory fp32 atomics are CAS loops on sm_100 and a warp-wide reduction of 32
// accumulators costs ~120 instructions, so the kernel is organised to make reductions rare:
//
//   unit    = (scene, 16 Morton-adjacent queries, 512 Morton-adjacent keys) = 8192 pairs, resident in shared memory
//   phase A   per pair, ONCE for all 8 vertices: the 6 axis transforms of an axis-aligned box (x+,x-,y+,y-,z+,z-)
//             -> a 16-byte record (6 x 16-bit interpolation fractions, 6 x 4-bit cell indices) + dS of the 4 heads
//             as the scaled fp16 the backward kernel already wrote for the dQ/dK GEMMs (8 bytes);
//   per vertex (bins = table cells, independent of query and key, so the whole unit sorts together):
//     B1/B2/B3  histogram, scan, counting-sort of the pair ids by cell (run-aggregated integer atomics);
//     B4        every warp walks a contiguous chunk of the sorted list, 32 pairs per step.  Inside a segment every
//               lane FFMA2s its 8 corners x 4 heads into 32 private registers; when the cell changes the 32 x 32
//               partial sums are transposed/reduced with shuffles and sent with ONE RED.ADD.F32 per lane to a
//               per-CTA private copy of the (zero-padded, (n+2)^3) table in global memory (L2 resident).  Segments
//               shorter than TINY pairs skip the registers: lane = (corner, head), pairs broadcast one by one.
//   a last kernel sums the private copies, drops the padding cells and undoes the fp16 gradient scale.
//
// Boxes that are not axis aligned (rotated `object_coords` boxes, arbitrary vertex sets) use the same records: their
// 3 transforms are recomputed per vertex into the slots that vertex reads (phase B0).
#include "rpe_internal.h"
#include "rpe_fast.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace dt3 {

#ifndef VDETR_DT_THREADS
#define VDETR_DT_THREADS 512
#define VDETR_DT_QB 16
#endif
constexpr int QB = VDETR_DT_QB, KC = 512, NP = QB * KC;       // pairs per unit
constexpr int THREADS = VDETR_DT_THREADS, WARPS = THREADS / 32;
constexpr int ITEMS = NP / THREADS;                  // pairs per thread in the per-pair phases (thread-strided)
static_assert(NP % THREADS == 0 && (NP / 4) % THREADS == 0 && KC % 128 == 0 && NP <= 8192, "unit shape");
#ifndef VDETR_DT_TINY
#define VDETR_DT_TINY 8
#endif
constexpr int TINY = VDETR_DT_TINY;                              // segments shorter than this use the broadcast mode
constexpr int MAX_N = 10;                            // shared-memory budget (and the fused forward kernel) stop at 10 points per axis
constexpr unsigned FULL = 0xffffffffu;

struct Params {
  int B, nQ, nK, nQp, nKp, n, R, P3;      // R = n + 1 base cells per axis, P3 = n + 2 padded table points per axis
  int qblocks, kchunks, units, dense_scale;
  float log_scale, c1, c0;
  const float4* xyz4;                     // [B][nKp]
  const float4* geo;                      // [B][nQp][9]
  const __half* dsb;                      // [(b*nQp + q)*4 + h][nKp]   scale * dS
  const int* qperm;                       // [B][nQ] Morton order of the queries
  const unsigned* absmax_bits;            // -> scale
  float* priv;                            // [gridDim.x][8][P3^3][4]
  int4 cost;                              // per-segment cost model of the accumulation phase (see seg_cost)
  int only_slow;                          // dt3 as the companion of dt5: only queries whose box is not axis aligned
  const int* slow_count;                  // [1] number of such queries in the whole call (only_slow: 0 -> nothing to do)
  unsigned long long* phase_clocks;       // [8] optional (developer): cycles per phase summed over CTAs, else null
};

__device__ __forceinline__ float scale_of(unsigned bits, int dense) { return vdetr_dt_scale(bits, dense); }

// one axis: cell index + 1 in [0, n] (15 = both corners outside the table) and the 16-bit fraction
__device__ __forceinline__ void axis_rec(float d, float ls, float c1, float c0, int n, unsigned& nib, unsigned& frac) {
  float t = tc::lg2_approx(fmaf(fabsf(d), ls, 1.0f)) * c1;
  float ts = copysignf(t, d);
  ts = fminf(fmaxf(ts, -c0 - 1.5f), (float)n - c0 + 0.5f);
  const float p = ts + c0;
  const float r = (p - 0.5f) + rpe::MAGIC;
  const float f = p - (r - rpe::MAGIC);
  const int n0 = __float_as_int(r) - rpe::MAGIC_BITS;
  nib = ((unsigned)(n0 + 1) <= (unsigned)n) ? (unsigned)(n0 + 1) : 15u;
  frac = min(__float2uint_rn(fminf(fmaxf(f, 0.f), 1.f) * 65536.0f), 65535u);      // fraction in units of 2^-16
}

// vertex sign table: 0:(+,+,-) 1:(+,-,-) 2:(-,-,-) 3:(-,+,-) 4:(+,+,+) 5:(+,-,+) 6:(-,-,+) 7:(-,+,+)
// -> slot selector (0 = '+', 1 = '-') per axis
__device__ __forceinline__ void vertex_slots(int v, int& xs, int& ys, int& zs) {
  xs = ((v & 3) == 2 || (v & 3) == 3) ? 1 : 0;
  ys = ((v & 3) == 1 || (v & 3) == 2) ? 1 : 0;
  zs = (v < 4) ? 1 : 0;
}

struct Smem {
  uint4* recs;        // [NP]  x: fx+ | fx- << 16   y: fy+ | fy-   z: fz+ | fz-   w: nibbles x+,x-,y+,y-,z+,z- (4 bits each)
  uint2* dsv;         // [NP]  4 x fp16 scaled dS
  uint16_t* sorted;   // [NP]
  int* hist;          // [nbins]   counts, then scatter cursors      } this region (+ costp[nbins+1]) doubles as
  int* offs;          // [nbins+1] exclusive scan                    } the key xyz staging buffer during phase A
  float4* sxyz;       // [KC] (aliases hist/offs)
  float4* sgeo;       // [QB][2]
  int* sq;            // [QB] query index (or -1), [QB] slow flags, misc
};

__host__ __device__ inline size_t region_bytes(int n) {
  const size_t nbins = (size_t)(n + 1) * (n + 1) * (n + 1);
  size_t r = (3 * nbins + 2 + 3) / 4 * 16;
  return r < (size_t)KC * 16 ? (size_t)KC * 16 : r;
}
__host__ __device__ inline size_t smem_bytes(int n) {
  return (size_t)NP * 16 + (size_t)NP * 8 + (size_t)NP * 2 + region_bytes(n) + QB * 2 * 16 + 160 * 4;
}

// transposed reduction: on return lane j holds sum over lanes of acc[j]
__device__ __forceinline__ float transpose_reduce(float (&a)[32], int lane) {
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float send = up ? a[i] : a[i + 16], keep = up ? a[i + 16] : a[i];
      a[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? a[i] : a[i + 8], keep = up ? a[i + 8] : a[i];
      a[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? a[i] : a[i + 4], keep = up ? a[i + 4] : a[i];
      a[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? a[i] : a[i + 2], keep = up ? a[i + 2] : a[i];
      a[i] = keep + __shfl_xor_sync(FULL, send, 2);
    }
  }
  const bool up = lane & 1;
  const float send = up ? a[0] : a[1], keep = up ? a[1] : a[0];
  return keep + __shfl_xor_sync(FULL, send, 1);
}

__device__ __forceinline__ float2 ffma2(float w, float2 d, float2 c) {
  unsigned long long rd, rc, ra, rb;
  const float2 ww = make_float2(w, w);
  ra = *reinterpret_cast<const unsigned long long*>(&ww);
  rb = *reinterpret_cast<const unsigned long long*>(&d);
  rc = *reinterpret_cast<const unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

// rough instruction cost of accumulating a segment of c pairs (units of half warp-instructions): register mode pays
// ~4.5 per pair plus a flush, broadcast mode ~16 per pair
__device__ __forceinline__ int seg_cost(int c, const int4 k) { return c == 0 ? 0 : (c < TINY ? k.x * c + k.y : k.z * c + k.w); }

// Run-aggregated shared-memory counter updates.  Lanes with equal, adjacent `bin` form a run (a warp holds 32
// consecutive keys of one query, so there are few runs); the head of a run adds the run length.  The run structure
// found for the histogram is packed into one word per pair and reused by the scatter:
//   bits [0,12) bin + 1 (0 = pair outside the table)   [12,17) rank inside the run   [17,23) run length
__device__ __forceinline__ unsigned run_pack(int bin, int lane) {
  // (straight-line on purpose: a warp-uniform early exit for the single-run case stops the compiler from
  // interleaving the unrolled per-pair chains and made the histogram phase 50 % slower)
  const int prev = __shfl_up_sync(FULL, bin, 1);
  const bool head = lane == 0 || bin != prev;
  const unsigned heads = __ballot_sync(FULL, head);
  const unsigned below = heads & (FULL >> (31 - lane));
  const int hl = 31 - __clz(below);
  const unsigned above = (hl == 31) ? 0u : (heads & ~((2u << hl) - 1u));
  const int end = above ? (__ffs(above) - 1) : 32;
  return (unsigned)(bin + 1) | ((unsigned)(lane - hl) << 12) | ((unsigned)(end - hl) << 17);
}

__device__ __forceinline__ float2 fmul2(float w, float2 d) {
  unsigned long long rd, ra, rb;
  const float2 ww = make_float2(w, w);
  ra = *reinterpret_cast<const unsigned long long*>(&ww);
  rb = *reinterpret_cast<const unsigned long long*>(&d);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

__global__ void __launch_bounds__(THREADS, 1) rpe_dtables_kernel(const Params P) {
  extern __shared__ __align__(16) uint8_t sm[];
  Smem S;
  S.recs = reinterpret_cast<uint4*>(sm);
  S.dsv = reinterpret_cast<uint2*>(sm + (size_t)NP * 16);
  S.sorted = reinterpret_cast<uint16_t*>(sm + (size_t)NP * 24);
  uint8_t* region = sm + (size_t)NP * 26;
  const int nbins = P.R * P.R * P.R;
  S.hist = reinterpret_cast<int*>(region);
  S.offs = S.hist + nbins;
  S.sxyz = reinterpret_cast<float4*>(region);
  S.sgeo = reinterpret_cast<float4*>(region + region_bytes(P.n));
  S.sq = reinterpret_cast<int*>(S.sgeo + QB * 2);
  int* s_slow = S.sq + QB;          // [QB]
  int* s_misc = S.sq + 2 * QB;      // [2..2+2*WARPS) scan scratch
  int* s_start = s_misc + 2 + 2 * WARPS;   // [WARPS+1] first sorted entry of every warp's share
  int* s_warpclk = s_start + WARPS + 1;    // [WARPS] developer clocks
  int* costp = S.offs + nbins + 1;  // [nbins+1] exclusive scan of the per-cell cost estimate

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (P.only_slow && *P.slow_count == 0) return;
  long long t_prev = clock64();
  auto tick = [&](int phase) {      // call right after a __syncthreads(): time since the previous tick -> phase
    if (P.phase_clocks && tid == 0) {
      const long long t = clock64();
      atomicAdd(P.phase_clocks + phase, (unsigned long long)(t - t_prev));
      t_prev = t;
    }
  };
  const int cells_pad = P.P3 * P.P3 * P.P3;
  float* my_priv = P.priv + (size_t)blockIdx.x * 8 * cells_pad * 4;
  const int corner = lane >> 2, hsel = lane & 3;
  const int cz = corner >> 2, cy = (corner >> 1) & 1, cx = corner & 1;
  // lane-constant affine maps fraction -> weight of this lane's corner along each axis (broadcast mode)
  const float mzs = cz ? 1.f : -1.f, mzo = cz ? 0.f : 1.f;
  const float mys = cy ? 1.f : -1.f, myo = cy ? 0.f : 1.f;
  const float mxs = cx ? 1.f : -1.f, mxo = cx ? 0.f : 1.f;

  for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
    int r = u;
    const int kc = r % P.kchunks; r /= P.kchunks;
    const int qb = r % P.qblocks;
    const int b = r / P.qblocks;
    const int q0 = qb * QB, k0 = kc * KC;

    __syncthreads();                      // previous unit completely done with shared memory
    if (tid < QB) {
      const int qi = q0 + tid;
      int q = -1, slow = 0;
      float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
      if (qi < P.nQ) {
        q = __ldg(P.qperm + (size_t)b * P.nQ + qi);
        const float4* g = P.geo + ((size_t)b * P.nQp + q) * 9;
        hi = __ldg(g); lo = __ldg(g + 1);
        slow = __float_as_int(hi.w) == 0;
        if (P.only_slow && !slow) q = -1;                    // axis-aligned boxes are the dt5 kernel's
      }
      S.sgeo[tid * 2] = hi; S.sgeo[tid * 2 + 1] = lo;
      S.sq[tid] = q; s_slow[tid] = slow;
    }
    for (int i = tid; i < KC; i += THREADS) {
      float4 kx = make_float4(1e9f, 1e9f, 1e9f, 0.f);
      if (k0 + i < P.nK) kx = __ldg(P.xyz4 + (size_t)b * P.nKp + k0 + i);
      S.sxyz[i] = kx;
    }
    __syncthreads();
    bool any_slow = false;
#pragma unroll
    for (int ql = 0; ql < QB; ++ql) any_slow = any_slow || (s_slow[ql] != 0 && S.sq[ql] >= 0);
    if (P.only_slow && !any_slow) continue;                  // (block-uniform: read from shared memory after the barrier)

    // ---- phase A: records.  A thread takes 4 consecutive keys of one query at a time so that the scaled-fp16 dS of
    // the 4 heads arrives as four 8-byte loads (64 loads of 2 bytes per thread left the phase latency bound).
    {
      constexpr int KG = KC / 4, GI = NP / 4 / THREADS;
#pragma unroll 1
      for (int gi = 0; gi < GI; ++gi) {
        const int g = gi * THREADS + tid, ql = g / KG, kl0 = (g % KG) * 4;
        const int q = S.sq[ql];
        uint2 hd[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) hd[h] = make_uint2(0u, 0u);
        const bool any = q >= 0 && k0 + kl0 < P.nK;
        if (any) {
          const __half* dp = P.dsb + ((size_t)b * P.nQp + q) * 4 * P.nKp + k0 + kl0;
#pragma unroll
          for (int h = 0; h < 4; ++h) hd[h] = __ldg(reinterpret_cast<const uint2*>(dp + (size_t)h * P.nKp));
        }
        const bool fast = any && !s_slow[ql];
        const float4 hi = S.sgeo[ql * 2], lo = S.sgeo[ql * 2 + 1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kl = kl0 + j, p = ql * KC + kl;
          uint4 rec = make_uint4(0u, 0u, 0u, 0x00FFFFFFu);
          uint2 dv = make_uint2(0u, 0u);
          if (any && k0 + kl < P.nK) {
            const int sh = 16 * (j & 1);
            const unsigned w0 = (j < 2) ? hd[0].x : hd[0].y, w1 = (j < 2) ? hd[1].x : hd[1].y;
            const unsigned w2 = (j < 2) ? hd[2].x : hd[2].y, w3 = (j < 2) ? hd[3].x : hd[3].y;
            dv = make_uint2(((w0 >> sh) & 0xFFFFu) | (((w1 >> sh) & 0xFFFFu) << 16), ((w2 >> sh) & 0xFFFFu) | (((w3 >> sh) & 0xFFFFu) << 16));
            if (fast) {
              const float4 kx = S.sxyz[kl];
              unsigned nxp, nxm, nyp, nym, nzp, nzm, fxp, fxm, fyp, fym, fzp, fzm;
              axis_rec(hi.x - kx.x, P.log_scale, P.c1, P.c0, P.n, nxp, fxp);
              axis_rec(lo.x - kx.x, P.log_scale, P.c1, P.c0, P.n, nxm, fxm);
              axis_rec(hi.y - kx.y, P.log_scale, P.c1, P.c0, P.n, nyp, fyp);
              axis_rec(lo.y - kx.y, P.log_scale, P.c1, P.c0, P.n, nym, fym);
              axis_rec(hi.z - kx.z, P.log_scale, P.c1, P.c0, P.n, nzp, fzp);
              axis_rec(lo.z - kx.z, P.log_scale, P.c1, P.c0, P.n, nzm, fzm);
              rec.x = fxp | (fxm << 16); rec.y = fyp | (fym << 16); rec.z = fzp | (fzm << 16);
              rec.w = nxp | (nxm << 4) | (nyp << 8) | (nym << 12) | (nzp << 16) | (nzm << 20);
            }
          }
          S.recs[p] = rec;
          S.dsv[p] = dv;
        }
      }
    }
    __syncthreads();                      // records complete; the xyz staging area becomes hist / offs
    tick(0);

    for (int vert = 0; vert < 8; ++vert) {
      int xs, ys, zs;
      vertex_slots(vert, xs, ys, zs);
      const int shx = 16 * xs, shy = 16 * ys, shz = 16 * zs;                 // fraction shifts
      const int nbx = 4 * xs, nby = 8 + 4 * ys, nbz = 16 + 4 * zs;           // nibble shifts

      // ---- B0: boxes that are not axis aligned: this vertex's own 3 transforms go into the slots it reads
      if (any_slow) {
        for (int it = 0; it < ITEMS; ++it) {
          const int p = it * THREADS + tid, ql = p / KC, kl = p % KC;          // ql is warp-uniform
          if (!s_slow[ql] || S.sq[ql] < 0 || k0 + kl >= P.nK) continue;
          const float4* g = P.geo + ((size_t)b * P.nQp + S.sq[ql]) * 9;
          const float* vv = reinterpret_cast<const float*>(g + 2);
          const float4 rot = __ldg(g + 8);
          const float4 kx = __ldg(P.xyz4 + (size_t)b * P.nKp + k0 + kl);
          const float dx = __ldg(vv + vert * 3) - kx.x, dy = __ldg(vv + vert * 3 + 1) - kx.y, dz = __ldg(vv + vert * 3 + 2) - kx.z;
          const float tx = rot.x * dx - rot.y * dy, ty = rot.y * dx + rot.x * dy;
          unsigned nx, ny, nz, fx, fy, fz;
          axis_rec(tx, P.log_scale, P.c1, P.c0, P.n, nx, fx);
          axis_rec(ty, P.log_scale, P.c1, P.c0, P.n, ny, fy);
          axis_rec(dz, P.log_scale, P.c1, P.c0, P.n, nz, fz);
          uint4 rec = S.recs[p];
          rec.x = (rec.x & ~(0xFFFFu << shx)) | (fx << shx);
          rec.y = (rec.y & ~(0xFFFFu << shy)) | (fy << shy);
          rec.z = (rec.z & ~(0xFFFFu << shz)) | (fz << shz);
          rec.w = (rec.w & ~((15u << nbx) | (15u << nby) | (15u << nbz))) | (nx << nbx) | (ny << nby) | (nz << nbz);
          S.recs[p] = rec;
        }
      }
      for (int i = tid; i < nbins; i += THREADS) S.hist[i] = 0;
      __syncthreads();
      tick(1);

      // ---- B1: histogram of the cells
      unsigned myrun[ITEMS];
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const unsigned w = S.recs[it * THREADS + tid].w;
        const int nx = (w >> nbx) & 15, ny = (w >> nby) & 15, nz = (w >> nbz) & 15;
        int bin = (nz * P.R + ny) * P.R + nx;
        if (max(nx, max(ny, nz)) == 15) bin = -1;
        const unsigned r = run_pack(bin, lane);
        myrun[it] = r;
        if ((r & 0x1F000u) == 0u && bin >= 0) atomicAdd(S.hist + bin, (int)(r >> 17));        // run head
      }
      __syncthreads();
      tick(2);

      // ---- B2: exclusive scan of the histogram (all threads, contiguous slices of bins) and of a per-cell cost
      // estimate that balances the accumulation phase over the warps
      {
        const int per = (nbins + THREADS - 1) / THREADS;
        const int lo = tid * per, hi = min(nbins, lo + per);
        int mine = 0, minec = 0;
        for (int i = lo; i < hi; ++i) { const int c = S.hist[i]; mine += c; minec += seg_cost(c, P.cost); }
        int x = mine, xc = minec;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(FULL, x, o), yc = __shfl_up_sync(FULL, xc, o);
          if (lane >= o) { x += y; xc += yc; }
        }
        if (lane == 31) { s_misc[2 + warp] = x; s_misc[2 + WARPS + warp] = xc; }
        __syncthreads();
        int wbase = 0, wbasec = 0;
        for (int w = 0; w < warp; ++w) { wbase += s_misc[2 + w]; wbasec += s_misc[2 + WARPS + w]; }
        int run = wbase + x - mine, runc = wbasec + xc - minec;
        for (int i = lo; i < hi; ++i) {
          const int c = S.hist[i];
          S.hist[i] = run | ((c < TINY) ? (int)0x80000000 : 0);      // scatter cursor; bit 31 marks a tiny segment
          S.offs[i] = run; costp[i] = runc;
          run += c; runc += seg_cost(c, P.cost);
        }
        if (tid == THREADS - 1) { S.offs[nbins] = run; costp[nbins] = runc; }
      }
      __syncthreads();
      tick(3);

      // ---- B3: every warp finds where its equal-cost share of the sorted list starts, then the counting-sort scatter
      {
        const int totalc = costp[nbins];
        const int target = (int)(((long long)totalc * warp) / WARPS);
        int lo = 0, hi = nbins;                       // costp[lo] <= target < costp[hi] (if totalc > 0)
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (costp[mid] <= target) lo = mid; else hi = mid;
        }
        const int c = S.offs[lo + 1] - S.offs[lo], cb = costp[lo + 1] - costp[lo];
        int start = S.offs[lo];
        if (cb > 0) start += min(c, (int)(((long long)(target - costp[lo]) * c) / cb));
        if (lane == 0) { s_start[warp] = start; if (warp == 0) s_start[WARPS] = S.offs[nbins]; }
      }
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const unsigned r = myrun[it];
        const int bin = (int)(r & 0xFFFu) - 1, rank = (int)((r >> 12) & 31u);
        int base = 0;
        if (rank == 0 && bin >= 0) base = atomicAdd(S.hist + bin, (int)(r >> 17));
        base = __shfl_sync(FULL, base, lane - rank);
        if (bin >= 0) S.sorted[(base & 0x7fffffff) + rank] = (uint16_t)((it * THREADS + tid) | ((base >> 16) & 0x8000));
      }
      __syncthreads();
      tick(4);

      // ---- B4: accumulate.  Warp w owns sorted[c0, c1).
      {
        const long long t_b4 = P.phase_clocks ? clock64() : 0;
        const int c0 = s_start[warp], c1 = s_start[warp + 1];
        float* tab = my_priv + (size_t)vert * cells_pad * 4;
        float2 acc[16];                   // [corner][head pair]
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = make_float2(0.f, 0.f);
        unsigned cur_key = 0xFFFFFFFEu;   // masked nibble word of the cell being accumulated (none yet)

        auto cell_addr = [&](unsigned key) -> float* {
          const int nx = (key >> nbx) & 15, ny = (key >> nby) & 15, nz = (key >> nbz) & 15;
          return tab + ((((nz + cz) * P.P3 + (ny + cy)) * P.P3 + (nx + cx)) << 2) + hsel;
        };
        auto flush = [&](unsigned key) {
          float a[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) { a[2 * j] = acc[j].x; a[2 * j + 1] = acc[j].y; acc[j] = make_float2(0.f, 0.f); }
          // a[corner * 4 + head] with head pairs (0,1) (2,3): index = corner*4 + hp*2 + e == lane's (corner, hsel)
          const float tot = transpose_reduce(a, lane);
          if (tot != 0.f) atomicAdd(cell_addr(key), tot);
        };

        // software pipeline: the entry / record / dS of step s+1 are loaded while step s is accumulated
        // (reads past c1 stay inside the shared-memory allocation and are masked to valid record ids)
        const unsigned vmask = (15u << nbx) | (15u << nby) | (15u << nbz);
        const unsigned selx = xs ? 0x7432u : 0x7410u, sely = ys ? 0x7432u : 0x7410u, selz = zs ? 0x7432u : 0x7410u;
        const uint16_t* sp = S.sorted + c0 + lane;
        unsigned ent_n = *sp;
        uint4 rec_n = S.recs[ent_n & 0x1FFFu];
        uint2 dv_n = S.dsv[ent_n & 0x1FFFu];
        for (int p0 = c0; p0 < c1; p0 += 32) {
          const bool live = p0 + lane < c1;
          const unsigned ent = ent_n;
          const uint4 rec = rec_n;
          const uint2 dv = dv_n;
          sp += 32;
          ent_n = *sp;
          rec_n = S.recs[ent_n & 0x1FFFu];
          dv_n = S.dsv[ent_n & 0x1FFFu];
          // cell identity = the 3 nibbles this vertex reads, left in place (dead lanes: a value no cell can have)
          const unsigned key = live ? (rec.w & vmask) : 0xFFFFFFFFu;
          // fraction u / 65536 without I2F: bytes (u.lo, u.hi, 0x00, 0x4B) = 2^23 + u ;  (2^23 + u) * 2^-16 - 128
          const float fx = fmaf(__uint_as_float(__byte_perm(rec.x, 0x4B000000u, selx)), 0x1p-16f, -128.f);
          const float fy = fmaf(__uint_as_float(__byte_perm(rec.y, 0x4B000000u, sely)), 0x1p-16f, -128.f);
          const float fz = fmaf(__uint_as_float(__byte_perm(rec.z, 0x4B000000u, selz)), 0x1p-16f, -128.f);
          float w[8];
          {
            const float2 wy = make_float2(1.f - fy, fy), wx = make_float2(1.f - fx, fx);
            const float2 a0 = fmul2(1.f - fz, wy), a1 = fmul2(fz, wy);          // (z0y0, z0y1) (z1y0, z1y1)
            const float2 w01 = fmul2(a0.x, wx), w23 = fmul2(a0.y, wx), w45 = fmul2(a1.x, wx), w67 = fmul2(a1.y, wx);
            w[0] = w01.x; w[1] = w01.y; w[2] = w23.x; w[3] = w23.y; w[4] = w45.x; w[5] = w45.y; w[6] = w67.x; w[7] = w67.y;
          }
          const float2 d01 = __half22float2(*reinterpret_cast<const __half2*>(&dv.x));
          const float2 d23 = __half22float2(*reinterpret_cast<const __half2*>(&dv.y));

          if (__all_sync(FULL, key == cur_key)) {          // the common step: 32 more pairs of the current cell
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              acc[2 * c] = ffma2(w[c], d01, acc[2 * c]);
              acc[2 * c + 1] = ffma2(w[c], d23, acc[2 * c + 1]);
            }
            continue;
          }
          unsigned todo = __ballot_sync(FULL, live);
          while (todo) {
            const int leader = __ffs(todo) - 1;
            const unsigned ksel = __shfl_sync(FULL, key, leader);
            const bool tiny = (__shfl_sync(FULL, ent, leader) & 0x8000u) != 0u;
            const unsigned grp = __ballot_sync(FULL, key == ksel);
            todo &= ~grp;
            if (ksel != cur_key && tiny) {
              // broadcast mode: lane = (corner, head); the pairs of the run are visited one by one
              float t = 0.f;
              unsigned g2 = grp;
              while (g2) {
                const int e = __ffs(g2) - 1;
                g2 &= g2 - 1;
                const float ez = __shfl_sync(FULL, fz, e), ey = __shfl_sync(FULL, fy, e), ex = __shfl_sync(FULL, fx, e);
                const unsigned dlo = __shfl_sync(FULL, dv.x, e), dhi = __shfl_sync(FULL, dv.y, e);
                const unsigned word = (hsel & 2) ? dhi : dlo;
                const float dval = __half2float(__ushort_as_half((unsigned short)((hsel & 1) ? (word >> 16) : (word & 0xFFFFu))));
                const float wz = fmaf(ez, mzs, mzo), wy = fmaf(ey, mys, myo), wx = fmaf(ex, mxs, mxo);
                t = fmaf(wz * wy * wx, dval, t);
              }
              if (t != 0.f) atomicAdd(cell_addr(ksel), t);
              continue;
            }
            if (ksel != cur_key) {
              if (cur_key != 0xFFFFFFFEu) flush(cur_key);
              cur_key = ksel;
            }
            if ((grp >> lane) & 1u) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                acc[2 * c] = ffma2(w[c], d01, acc[2 * c]);
                acc[2 * c + 1] = ffma2(w[c], d23, acc[2 * c + 1]);
              }
            }
          }
        }
        if (cur_key != 0xFFFFFFFEu) flush(cur_key);
        if (P.phase_clocks && lane == 0) s_warpclk[warp] = (int)(clock64() - t_b4);
      }
      __syncthreads();                    // hist / offs / sorted are rewritten by the next vertex
      tick(5);
      if (P.phase_clocks && tid == 0) {         // balance of the accumulation phase: mean and max warp time
        int mx = 0, sum = 0;
        for (int w = 0; w < WARPS; ++w) { mx = max(mx, s_warpclk[w]); sum += s_warpclk[w]; }
        atomicAdd(P.phase_clocks + 6, (unsigned long long)(sum / WARPS));
        atomicAdd(P.phase_clocks + 7, (unsigned long long)mx);
      }
    }
  }
}


}  // namespace dt3

// =====================================================================================================================
// dt5 (opt-in: VDETR_DT_IMPL=5): the same adjoint for AXIS-ALIGNED boxes with the accumulation done by warp-level
// tensor-core MMAs (mma.sync.m16n8k16 + ldmatrix) instead of per-lane register accumulators + shuffle transposes.
// Measured on B200 at 8 x 1024 x 4096: 3.8 ms against dt3's 3.7 ms -- see DESIGN.md 4.3 for why it does not win: staging
// the fp16 weight fragments through shared memory costs as many instructions as the FFMA2s it replaces, and the sort
// (halved here) was the smaller half of dt3.  Kept as a tested alternative.
//
//   unit      = (scene, 12 Morton-adjacent queries, 512 Morton-adjacent keys) = 6144 pairs; phase A as in dt3.
//   round     = a PAIR of vertices that differ only in the x sign (0,3) (1,2) (4,7) (5,6): they share the y and z
//               transforms, hence the (z, y) part of their table cell and 4 of the 6 weight factors: ONE counting sort
//               per round, by the cell of vertex A (x+).
//   vertex B (x-) accumulates into an x WINDOW: its 16 MMA rows are (4 (z,y) corners) x (4 consecutive x points).  Inside
//               one cell of A the cell of B is x_A - delta with delta in {0,1,2} for all but a handful of pairs (class 1),
//               so with the window based at x_A - 2 vertex B never changes accumulators inside A's segment.
//   layout    : class-0 bins are laid out first, each padded with dummy pairs to a multiple of 16, so that every 16-pair
//               MMA block of that region belongs to exactly one cell; class-1 pairs follow unpadded and go through a
//               masked path with B's window based at its own cell.
//   accumulate: lane = pair computes the 8 corner weights of A and the 16 window weights of B (fp32, rounded once to
//               fp16) and leaves them with the 4 heads' dS in a per-warp staging tile; per 16 pairs ldmatrix builds the
//               fragments of two MMAs  D[16 x 8] += W^T[16 x 16 pairs] * dS[16 pairs x 8].  D stays in registers while
//               the cell does not change; a cell change costs one vector RED per lane (keys are linear table offsets).
//               Work is drawn in chunks from a shared counter, the expensive unpadded end of the list first.
// Queries whose box is not axis aligned are skipped here (zero weights) and handled by dt3 with only_slow = 1.
namespace dt5 {

using dt3::axis_rec;
using dt3::FULL;
using dt3::Params;

#ifndef VDETR_DT5_CHUNK
#define VDETR_DT5_CHUNK 4
#endif

// run structure of a warp's 32 bins (see dt3::run_pack); bins need 13 bits here:
//   bits [0,13) bin + 1 (0 = pair outside the table)   [13,18) rank inside the run   [18,24) run length
__device__ __forceinline__ unsigned run_pack13(int bin, int lane) {
  const int prev = __shfl_up_sync(FULL, bin, 1);
  const bool head = lane == 0 || bin != prev;
  const unsigned heads = __ballot_sync(FULL, head);
  const unsigned below = heads & (FULL >> (31 - lane));
  const int hl = 31 - __clz(below);
  const unsigned above = (hl == 31) ? 0u : (heads & ~((2u << hl) - 1u));
  const int end = above ? (__ffs(above) - 1) : 32;
  return (unsigned)(bin + 1) | ((unsigned)(lane - hl) << 13) | ((unsigned)(end - hl) << 18);
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void red_add_v2(float* addr, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(x), "f"(y) : "memory");
}
// 16-bit lanes of a fragment register kept where the membership bit of their pair is set (bit 0 -> low half, bit 1 -> high)
__device__ __forceinline__ uint32_t keep_halves(uint32_t v, unsigned bits) {
  const uint32_t m = ((0u - (bits & 1u)) & 0x0000FFFFu) | ((0u - ((bits >> 1) & 1u)) & 0xFFFF0000u);
  return v & m;
}

constexpr int QB = 12, KC = 512, NP = QB * KC;
constexpr int THREADS = 512, WARPS = 16, ITEMS = NP / THREADS;
constexpr int XI = 12;
constexpr int SORT_CAP = NP + (NP / 16) * 15 + 128;      // large bins padded to 16 (at most NP/16 of them) + trailing dummies / prefetch
constexpr int STAGE_BYTES = 4 * 32 * 16;                 // per warp: A weights | B window halves 0-7 | 8-15 | dS  (16 B per pair each)
constexpr int CHUNK = VDETR_DT5_CHUNK;                   // steps per unit of dynamically scheduled accumulate work
constexpr unsigned WILD = 0xFFFFFFFFu;
static_assert((SORT_CAP * 2) % 16 == 0 && ((NP + 2) * 24) % 16 == 0, "16-byte alignment of the shared-memory regions");

__host__ __device__ inline int nbins_of(int n) { return (n + 1) * (n + 1) * XI * 2; }
__host__ __device__ inline size_t region_bytes(int n) {
  const size_t r = ((size_t)nbins_of(n) + 1 + 3) / 4 * 16;
  return r < (size_t)KC * 16 ? (size_t)KC * 16 : r;
}
__host__ __device__ inline size_t smem_bytes(int n) {
  return (size_t)(NP + 2) * 24 + (size_t)SORT_CAP * 2 + region_bytes(n) + WARPS * STAGE_BYTES + QB * 2 * 16 + 64 * 4 + 128;
}

__global__ void __launch_bounds__(THREADS, 1) rpe_dtables_win_kernel(const Params P) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* s_recs = reinterpret_cast<uint4*>(sm);                                     // [NP + 2], entry NP = dummy
  uint2* s_dsv = reinterpret_cast<uint2*>(sm + (size_t)(NP + 2) * 16);
  uint16_t* s_sorted = reinterpret_cast<uint16_t*>(sm + (size_t)(NP + 2) * 24);     // [SORT_CAP]
  uint8_t* region = sm + (size_t)(NP + 2) * 24 + (size_t)SORT_CAP * 2;
  int* s_hist = reinterpret_cast<int*>(region);                                     // [nbins + 1] counts -> cursors
  float4* s_xyz = reinterpret_cast<float4*>(region);                                // [KC] phase A only (aliases hist)
  uint8_t* s_stage = region + region_bytes(P.n);
  float4* s_geo = reinterpret_cast<float4*>(s_stage + WARPS * STAGE_BYTES);         // [QB][2]
  int* s_q = reinterpret_cast<int*>(s_geo + QB * 2);                                // [QB]
  int* s_misc = s_q + 16;      // [0] end of the sorted list  [1] chunk counter  [2] end of the padded region  [4..4+4*WARPS) scan scratch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbins = nbins_of(P.n);
  const int R = P.R;
  const int cells_pad = P.P3 * P.P3 * P.P3;
  float* my_priv = P.priv + (size_t)blockIdx.x * 8 * cells_pad * 4;
  long long t_prev = clock64();
  auto tick = [&](int phase) {
    if (P.phase_clocks && tid == 0) {
      const long long t = clock64();
      atomicAdd(P.phase_clocks + phase, (unsigned long long)(t - t_prev));
      t_prev = t;
    }
  };
  const int fg = lane >> 2, ft = lane & 3;                        // fragment row / column pair of this lane
  uint8_t* my_stage = s_stage + warp * STAGE_BYTES;
  const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(my_stage);
  // ldmatrix row addresses (+ 256 for the second 16-pair block of a step)
  const uint32_t rowA = stage_u32 + (((lane >> 3) & 1) * 8 + (lane & 7)) * 16;                                      // x2: WA k0-7, k8-15
  const uint32_t rowB = stage_u32 + 512 + ((lane >> 3) & 1) * 512 + (((lane >> 4) & 1) * 8 + (lane & 7)) * 16;     // x4: WB0 k0-7, WB1 k0-7, WB0 k8-15, WB1 k8-15
  const uint32_t rowS = stage_u32 + 1536 + (((lane >> 3) & 1) * 8 + (lane & 7)) * 16;                              // x2: dS k0-7, k8-15

  if (tid == 0) {
    s_recs[NP] = make_uint4(0u, 0u, 0u, 0x000000FFu);
    s_dsv[NP] = make_uint2(0u, 0u);
  }

  for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
    int r = u;
    const int kc = r % P.kchunks; r /= P.kchunks;
    const int qb = r % P.qblocks;
    const int b = r / P.qblocks;
    const int q0 = qb * QB, k0 = kc * KC;

    __syncthreads();
    if (tid < QB) {
      const int qi = q0 + tid;
      int q = -1;
      float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
      if (qi < P.nQ) {
        q = __ldg(P.qperm + (size_t)b * P.nQ + qi);
        const float4* g = P.geo + ((size_t)b * P.nQp + q) * 9;
        hi = __ldg(g); lo = __ldg(g + 1);
        if (__float_as_int(hi.w) == 0) q = -1;                    // not axis aligned: handled by the dt3 kernel
      }
      s_geo[tid * 2] = hi; s_geo[tid * 2 + 1] = lo;
      s_q[tid] = q;
    }
    for (int i = tid; i < KC; i += THREADS) {
      float4 kx = make_float4(1e9f, 1e9f, 1e9f, 0.f);
      if (k0 + i < P.nK) kx = __ldg(P.xyz4 + (size_t)b * P.nKp + k0 + i);
      s_xyz[i] = kx;
    }
    __syncthreads();

    // ---- phase A: records + dS (as in dt3)
    {
      constexpr int KG = KC / 4, GI = NP / 4 / THREADS;
#pragma unroll 1
      for (int gi = 0; gi < GI; ++gi) {
        const int g = gi * THREADS + tid, ql = g / KG, kl0 = (g % KG) * 4;
        const int q = s_q[ql];
        uint2 hd[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) hd[h] = make_uint2(0u, 0u);
        const bool any = q >= 0 && k0 + kl0 < P.nK;
        if (any) {
          const __half* dp = P.dsb + ((size_t)b * P.nQp + q) * 4 * P.nKp + k0 + kl0;
#pragma unroll
          for (int h = 0; h < 4; ++h) hd[h] = __ldg(reinterpret_cast<const uint2*>(dp + (size_t)h * P.nKp));
        }
        const float4 hi = s_geo[ql * 2], lo = s_geo[ql * 2 + 1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kl = kl0 + j, p = ql * KC + kl;
          uint4 rec = make_uint4(0u, 0u, 0u, 0x00FFFFFFu);
          uint2 dv = make_uint2(0u, 0u);
          if (any && k0 + kl < P.nK) {
            const int sh = 16 * (j & 1);
            const unsigned w0 = (j < 2) ? hd[0].x : hd[0].y, w1 = (j < 2) ? hd[1].x : hd[1].y;
            const unsigned w2 = (j < 2) ? hd[2].x : hd[2].y, w3 = (j < 2) ? hd[3].x : hd[3].y;
            dv = make_uint2(((w0 >> sh) & 0xFFFFu) | (((w1 >> sh) & 0xFFFFu) << 16), ((w2 >> sh) & 0xFFFFu) | (((w3 >> sh) & 0xFFFFu) << 16));
            const float4 kx = s_xyz[kl];
            unsigned nxp, nxm, nyp, nym, nzp, nzm, fxp, fxm, fyp, fym, fzp, fzm;
            axis_rec(hi.x - kx.x, P.log_scale, P.c1, P.c0, P.n, nxp, fxp);
            axis_rec(lo.x - kx.x, P.log_scale, P.c1, P.c0, P.n, nxm, fxm);
            axis_rec(hi.y - kx.y, P.log_scale, P.c1, P.c0, P.n, nyp, fyp);
            axis_rec(lo.y - kx.y, P.log_scale, P.c1, P.c0, P.n, nym, fym);
            axis_rec(hi.z - kx.z, P.log_scale, P.c1, P.c0, P.n, nzp, fzp);
            axis_rec(lo.z - kx.z, P.log_scale, P.c1, P.c0, P.n, nzm, fzm);
            rec.x = fxp | (fxm << 16); rec.y = fyp | (fym << 16); rec.z = fzp | (fzm << 16);
            rec.w = nxp | (nxm << 4) | (nyp << 8) | (nym << 12) | (nzp << 16) | (nzm << 20);
          }
          s_recs[p] = rec;
          s_dsv[p] = dv;
        }
      }
    }
    __syncthreads();
    tick(0);

    for (int round = 0; round < 4; ++round) {
      const int ys = round & 1, zs = (round < 2) ? 1 : 0;
      const int vertA = (round < 2 ? 0 : 4) + ys, vertB = (round < 2 ? 3 : 7) - ys;
      const int nby = 8 + 4 * ys, nbz = 16 + 4 * zs;

      for (int i = tid; i <= nbins; i += THREADS) s_hist[i] = 0;
      {                                   // the whole sorted array starts as dummy entries (padding of the large bins)
        const uint32_t dd = (uint32_t)NP | ((uint32_t)NP << 16);
        uint4* so = reinterpret_cast<uint4*>(s_sorted);
        for (int i = tid; i < SORT_CAP / 8; i += THREADS) so[i] = make_uint4(dd, dd, dd, dd);
      }
      __syncthreads();
      tick(1);

      // ---- B1: histogram.  bin = (z, y, x_A | 11, class); class 0: A inside the table and B inside A's window (or outside
      // the table: zero weights); class 1: everything else that touches the table.
      unsigned myrun[ITEMS];
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const unsigned w = s_recs[it * THREADS + tid].w;
        const int xa = w & 15, xb = (w >> 4) & 15, ny = (w >> nby) & 15, nz = (w >> nbz) & 15;
        const bool a_ok = xa != 15, b_ok = xb != 15;
        int bin = -1;
        if (max(ny, nz) != 15 && (a_ok || b_ok)) {
          const int cls = (a_ok && (!b_ok || (unsigned)(xa - xb) <= 2u)) ? 0 : 1;
          bin = (((nz * R + ny) * XI + (a_ok ? xa : XI - 1)) << 1) | cls;
        }
        const unsigned rp = run_pack13(bin, lane);
        myrun[it] = rp;
        if ((rp & 0x3E000u) == 0u && bin >= 0) atomicAdd(s_hist + bin, (int)(rp >> 18));
      }
      __syncthreads();
      tick(2);

      // ---- B2: two exclusive scans in one pass: class-0 bins with >= th pairs, each padded to a multiple of 16, first; the
      // other bins behind them.  th = 1 (every class-0 bin padded: no block of the padded region ever straddles two cells)
      // unless the padding would overflow the sorted array, then 16 (always fits: at most NP / 16 such bins).
      {
        const int per = (nbins + THREADS - 1) / THREADS;
        const int lo = tid * per, hi = min(nbins, lo + per);
        // per thread: padded size of the front region and pair count of the rest, for th = 1 and for th = 16
        int mL1 = 0, mS1 = 0, mL16 = 0, mS16 = 0;
        for (int i = lo; i < hi; ++i) {
          const int c = s_hist[i];
          const int a = (c + 15) & ~15;
          if (!(i & 1)) { mL1 += a; if (c >= 16) mL16 += a; else mS16 += c; }
          else { mS1 += c; mS16 += c; }
        }
        int xL1 = mL1, xS1 = mS1, xL16 = mL16, xS16 = mS16;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int a = __shfl_up_sync(FULL, xL1, o), b2 = __shfl_up_sync(FULL, xS1, o);
          const int c2 = __shfl_up_sync(FULL, xL16, o), d2 = __shfl_up_sync(FULL, xS16, o);
          if (lane >= o) { xL1 += a; xS1 += b2; xL16 += c2; xS16 += d2; }
        }
        if (lane == 31) { s_misc[4 + warp] = xL1; s_misc[4 + WARPS + warp] = xS1; s_misc[4 + 2 * WARPS + warp] = xL16; s_misc[4 + 3 * WARPS + warp] = xS16; }
        __syncthreads();
        int bL1 = 0, bS1 = 0, bL16 = 0, bS16 = 0, tL1 = 0, tS1 = 0, tL16 = 0;
        for (int w = 0; w < WARPS; ++w) {
          const int a = s_misc[4 + w], b2 = s_misc[4 + WARPS + w], c2 = s_misc[4 + 2 * WARPS + w], d2 = s_misc[4 + 3 * WARPS + w];
          if (w < warp) { bL1 += a; bS1 += b2; bL16 += c2; bS16 += d2; }
          tL1 += a; tS1 += b2; tL16 += c2;
        }
        // padded region (rounded up to a whole step) + unpadded remainder + trailing dummies / prefetch must fit
        const int th = (((tL1 + 31) & ~31) + tS1 + 96 <= SORT_CAP) ? 1 : 16;
        const int mineL = th == 1 ? mL1 : mL16, mineS = th == 1 ? mS1 : mS16;
        const int xL = th == 1 ? xL1 : xL16, xS = th == 1 ? xS1 : xS16;
        const int baseL = th == 1 ? bL1 : bL16, baseS = th == 1 ? bS1 : bS16, totL = th == 1 ? tL1 : tL16;
        const int large_end = (totL + 31) & ~31;               // the masked region starts on a step boundary
        int runL = baseL + xL - mineL, runS = large_end + baseS + xS - mineS;
        for (int i = lo; i < hi; ++i) {
          const int c = s_hist[i];
          if (c >= th && !(i & 1)) { s_hist[i] = runL; runL += (c + 15) & ~15; }
          else { s_hist[i] = runS; runS += c; }
        }
        if (tid == THREADS - 1) { s_misc[0] = runS; s_misc[2] = large_end; s_misc[1] = 0; }
      }
      __syncthreads();
      tick(3);

      // ---- B3: scatter
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const unsigned rp = myrun[it];
        const int bin = (int)(rp & 0x1FFFu) - 1, rank = (int)((rp >> 13) & 31u);
        int base = 0;
        if (rank == 0 && bin >= 0) base = atomicAdd(s_hist + bin, (int)(rp >> 18));
        base = __shfl_sync(FULL, base, lane - rank);
        if (bin >= 0) s_sorted[base + rank] = (uint16_t)(it * THREADS + tid);
      }
      __syncthreads();
      tick(4);

      // ---- B4: accumulate
      {
        const int list_end = s_misc[0], large_steps = s_misc[2] >> 5;
        const int steps = (list_end + 31) >> 5;
        float* tabA = my_priv + (size_t)vertA * cells_pad * 4;
        float* tabB = my_priv + (size_t)vertB * cells_pad * 4;
        float dA[4] = {0.f, 0.f, 0.f, 0.f};        // A: corner fg, heads 2ft, 2ft+1 ([2], [3]: the unused rows 8-15, always 0)
        float dB[4] = {0.f, 0.f, 0.f, 0.f};        // B window: rows fg and fg + 8 = ((z,y) corner, x slot)
        // Cell keys are LINEAR table offsets ((nz * P3 + ny) * P3 + x, in cells of the zero-padded private table; for B: x =
        // window base + 2), so a flush is one add per lane: element = key * 4 + a lane-constant offset of this lane's row.
        // A fragment element that maps outside the table (window slots left of x = 0 / right of the padding, vertices
        // outside the table) never receives a weight, so "value != 0" is also the bounds check.
        unsigned curA = WILD, curB = WILD;
        const int offA = ((((fg >> 2) * P.P3 + ((fg >> 1) & 1)) * P.P3 + (fg & 1)) << 2) + 2 * ft;
        const int offB = (((((fg >> 2) & 1) * P.P3) + (fg & 3) - 2) << 2) + 2 * ft;       // rows fg: z corner 0, y corner fg >> 2
        const int offB2 = (P.P3 * P.P3) << 2;                                              // rows fg + 8: z corner 1

        auto flushA = [&]() {
          if (curA != WILD && ft < 2 && (dA[0] != 0.f || dA[1] != 0.f)) red_add_v2(tabA + (int)(curA << 2) + offA, dA[0], dA[1]);
          dA[0] = 0.f; dA[1] = 0.f;
        };
        auto flushB = [&]() {
          if (curB != WILD && ft < 2) {
            float* a0 = tabB + (int)(curB << 2) + offB;
            if (dB[0] != 0.f || dB[1] != 0.f) red_add_v2(a0, dB[0], dB[1]);
            if (dB[2] != 0.f || dB[3] != 0.f) red_add_v2(a0 + offB2, dB[2], dB[3]);
          }
          dB[0] = 0.f; dB[1] = 0.f; dB[2] = 0.f; dB[3] = 0.f;
        };
        auto mmaA = [&](uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) { mma_16816(dA, a0, 0u, a2, 0u, b0, b1); };

        const unsigned sely = ys ? 0x7432u : 0x7410u, selz = zs ? 0x7432u : 0x7410u;
        for (;;) {
          int chunk = 0;
          if (lane == 0) chunk = atomicAdd(s_misc + 1, 1);
          chunk = __shfl_sync(FULL, chunk, 0);
          // chunks are handed out from the END of the list: the unpadded tail (many cells per step) is the expensive part and
          // must not be what the last warps are still working on while the others wait at the barrier
          const int s1 = steps - chunk * CHUNK;
          if (s1 <= 0) break;
          const int s0 = max(0, s1 - CHUNK);
          const uint16_t* sp = s_sorted + s0 * 32 + lane;
          unsigned ent_n = *sp;
          uint4 rec_n = s_recs[ent_n];
          uint2 dv_n = s_dsv[ent_n];
          for (int st = s0; st < s1; ++st) {
            const unsigned ent = ent_n;
            const uint4 rec = rec_n;
            const uint2 dv = dv_n;
            sp += 32;
            ent_n = *sp;
            rec_n = s_recs[ent_n];
            dv_n = s_dsv[ent_n];
            const bool tail = st >= large_steps;             // warp-uniform: masked region, B's window based at its own cell

            const unsigned xa = rec.w & 15u, xb = (rec.w >> 4) & 15u;
            const unsigned zy = (((rec.w >> nbz) & 15u) * (unsigned)P.P3 + ((rec.w >> nby) & 15u)) * (unsigned)P.P3;
            // (a vertex outside the table has zero weights: its key only has to be a harmless in-range cell)
            const unsigned xa1 = xa == 15u ? 0u : xa;
            unsigned keyA = zy + xa1, keyB = zy + (tail ? (xb == 15u ? 2u : xb + 2u) : xa1);
            if (ent == (unsigned)NP) { keyA = WILD; keyB = WILD; }
            const unsigned slot = (tail || xb == 15u) ? 0u : (2u - (xa - xb));        // class 0 in the large region: xa - xb in {0,1,2}
            const float fz = fmaf(__uint_as_float(__byte_perm(rec.z, 0x4B000000u, selz)), 0x1p-16f, -128.f);
            const float fy = fmaf(__uint_as_float(__byte_perm(rec.y, 0x4B000000u, sely)), 0x1p-16f, -128.f);
            const float fa = fmaf(__uint_as_float(__byte_perm(rec.x, 0x4B000000u, 0x7410u)), 0x1p-16f, -128.f);
            const float fb = fmaf(__uint_as_float(__byte_perm(rec.x, 0x4B000000u, 0x7432u)), 0x1p-16f, -128.f);
            {
              const float2 wy = make_float2(1.f - fy, fy);
              const float2 z0 = dt3::fmul2(1.f - fz, wy), z1 = dt3::fmul2(fz, wy);
              const float2 wa = xa == 15u ? make_float2(0.f, 0.f) : make_float2(1.f - fa, fa);
              const float2 wb = xb == 15u ? make_float2(0.f, 0.f) : make_float2(1.f - fb, fb);
              const float2 a01 = dt3::fmul2(z0.x, wa), a23 = dt3::fmul2(z0.y, wa), a45 = dt3::fmul2(z1.x, wa), a67 = dt3::fmul2(z1.y, wa);
              const float2 b0 = dt3::fmul2(z0.x, wb), b1 = dt3::fmul2(z0.y, wb), b2 = dt3::fmul2(z1.x, wb), b3 = dt3::fmul2(z1.y, wb);
              // B window row: halves [(z,y) corner j][x slot 0..3], the pair (x0, x1) of corner j sits at slots (slot, slot + 1)
              const unsigned sh = 16u * slot;
              const unsigned long long v0 = (unsigned long long)tc::pack_f16x2(b0.x, b0.y) << sh, v1 = (unsigned long long)tc::pack_f16x2(b1.x, b1.y) << sh;
              const unsigned long long v2 = (unsigned long long)tc::pack_f16x2(b2.x, b2.y) << sh, v3 = (unsigned long long)tc::pack_f16x2(b3.x, b3.y) << sh;
              __syncwarp();                                 // the previous step's ldmatrix reads are complete
              uint4* stg = reinterpret_cast<uint4*>(my_stage);
              stg[lane] = make_uint4(tc::pack_f16x2(a01.x, a01.y), tc::pack_f16x2(a23.x, a23.y), tc::pack_f16x2(a45.x, a45.y),
                                     tc::pack_f16x2(a67.x, a67.y));
              stg[32 + lane] = make_uint4((uint32_t)v0, (uint32_t)(v0 >> 32), (uint32_t)v1, (uint32_t)(v1 >> 32));
              stg[64 + lane] = make_uint4((uint32_t)v2, (uint32_t)(v2 >> 32), (uint32_t)v3, (uint32_t)(v3 >> 32));
              stg[96 + lane] = make_uint4(dv.x, dv.y, 0u, 0u);
              __syncwarp();
            }
            // per 16-pair block: is it one cell for A and one window for B?  (dummy pairs match anything.)  In the padded
            // region every block is, by construction; only the unpadded tail has to look.
            unsigned eqA = FULL, eqB = FULL;
            if (tail) {
              const unsigned refA = __shfl_sync(FULL, keyA, lane & 16), refB = __shfl_sync(FULL, keyB, lane & 16);
              eqA = __ballot_sync(FULL, keyA == refA || keyA == WILD);
              eqB = __ballot_sync(FULL, keyB == refB || keyB == WILD);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const unsigned kA = __shfl_sync(FULL, keyA, 16 * h);
              const unsigned kB = tail ? __shfl_sync(FULL, keyB, 16 * h) : kA;      // padded region: B's window hangs on A's cell
              if (kA == WILD && ((eqA >> (16 * h)) & 0xFFFFu) == 0xFFFFu) continue;          // a block of dummy pairs only
              uint32_t a0, a2, w0, w1, w2, w3, b0, b1;
              ldmatrix_x2_trans(rowA + h * 256, a0, a2);
              ldmatrix_x4_trans(rowB + h * 256, w0, w1, w2, w3);
              ldmatrix_x2_trans(rowS + h * 256, b0, b1);
              const bool uniA = kA != WILD && ((eqA >> (16 * h)) & 0xFFFFu) == 0xFFFFu;
              const bool uniB = kB != WILD && ((eqB >> (16 * h)) & 0xFFFFu) == 0xFFFFu;
              if (uniA && uniB) {
                if (kA != curA) { flushA(); curA = kA; }
                if (kB != curB) { flushB(); curB = kB; }
                mmaA(a0, a2, b0, b1);
                mma_16816(dB, w0, w1, w2, w3, b0, b1);
                continue;
              }
              // a block that straddles cells (only behind the large region): one masked MMA per distinct key and vertex
#pragma unroll 1
              for (int v = 0; v < 2; ++v) {
                const unsigned key = v ? keyB : keyA;
                unsigned todo = __ballot_sync(FULL, key != WILD) & (0xFFFFu << (16 * h));
                while (todo) {
                  const int leader = __ffs(todo) - 1;
                  const unsigned ksel = __shfl_sync(FULL, key, leader);
                  const unsigned grp = __ballot_sync(FULL, key == ksel) & (0xFFFFu << (16 * h));
                  todo &= ~grp;
                  const unsigned bits = (grp >> (16 * h)) >> (2 * ft);
                  if (v == 0) {
                    if (ksel != curA) { flushA(); curA = ksel; }
                    mmaA(keep_halves(a0, bits), keep_halves(a2, bits >> 8), b0, b1);
                  } else {
                    if (ksel != curB) { flushB(); curB = ksel; }
                    mma_16816(dB, keep_halves(w0, bits), keep_halves(w1, bits), keep_halves(w2, bits >> 8), keep_halves(w3, bits >> 8), b0, b1);
                  }
                }
              }
            }
          }
          flushA(); flushB();
          curA = WILD; curB = WILD;
        }
      }
      __syncthreads();
      tick(5);
    }
  }
}

}  // namespace dt5

namespace dt3 {

// sum of the per-CTA private tables, without the padding cells, times 1 / scale.  A CTA owns 32 consecutive output
// elements; its 8 warps split the copies (the loads of a warp are 128 contiguous bytes of one copy).
__global__ void __launch_bounds__(256) rpe_dtables_reduce_kernel(const float* __restrict__ priv, int copies, int n, int P3,
                                                                 float* __restrict__ out, const unsigned* absmax_bits, int dense) {
  __shared__ float part[8][32];
  const float inv = 1.0f / scale_of(*absmax_bits, dense);
  const int total = 8 * n * n * n * 4;
  const size_t copy_stride = (size_t)8 * P3 * P3 * P3 * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = blockIdx.x * 32; i0 < total; i0 += gridDim.x * 32) {
    const int i = i0 + lane;
    float s = 0.f;
    if (i < total) {
      const int h = i & 3;
      int r = i >> 2;
      const int x = r % n; r /= n;
      const int y = r % n; r /= n;
      const int z = r % n;
      const int v = r / n;
      const size_t src = ((((size_t)v * P3 + z + 1) * P3 + y + 1) * P3 + x + 1) * 4 + h;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int c = warp;
      for (; c + 24 < copies; c += 32) {
        s0 += priv[c * copy_stride + src]; s1 += priv[(c + 8) * copy_stride + src];
        s2 += priv[(c + 16) * copy_stride + src]; s3 += priv[(c + 24) * copy_stride + src];
      }
      for (; c < copies; c += 8) s0 += priv[c * copy_stride + src];
      s = (s0 + s1) + (s2 + s3);
    }
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][lane];
      out[i] = t * inv;
    }
    __syncthreads();
  }
}

// Morton order of the queries of every scene (box centre = mean of vertices 2 (-,-,-) and 4 (+,+,+)): neighbouring
// boxes see the keys in the same table cells, which halves the number of segments per unit.
constexpr int QSORT_MAX = 4096;
__global__ void __launch_bounds__(1024, 1) rpe_dtables_qorder_kernel(const float4* __restrict__ geo, int nQ, int nQp, int* __restrict__ qperm,
                                                                     int* __restrict__ slow_count) {
  __shared__ unsigned long long keys[QSORT_MAX];
  __shared__ float red[6][32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* out = qperm + (size_t)b * nQ;
  {                                       // number of queries whose box is not axis aligned (dt3 handles those)
    int ns = 0;
    for (int i = tid; i < nQ; i += blockDim.x) ns += __float_as_int(__ldg(&geo[((size_t)b * nQp + i) * 9].w)) == 0;
    ns = __reduce_add_sync(FULL, ns);
    if (lane == 0 && ns) atomicAdd(slow_count, ns);
  }
  if (nQ > QSORT_MAX) {                   // identity: still correct, only less coherent
    for (int i = tid; i < nQ; i += blockDim.x) out[i] = i;
    return;
  }
  int N = 1;
  while (N < nQ) N <<= 1;
  float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int i = tid; i < nQ; i += blockDim.x) {
    const float* v = reinterpret_cast<const float*>(geo + ((size_t)b * nQp + i) * 9 + 2);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float c = 0.5f * (v[6 + a] + v[12 + a]);
      if (c == c) { mn[a] = fminf(mn[a], c); mx[a] = fmaxf(mx[a], c); }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
    }
    if (lane == 0) { red[a][warp] = mn[a]; red[3 + a][warp] = mx[a]; }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float m0 = 3.0e38f, m1 = -3.0e38f;
    for (int w = 0; w < nw; ++w) { m0 = fminf(m0, red[a][w]); m1 = fmaxf(m1, red[3 + a][w]); }
    mn[a] = m0; mx[a] = m1;
  }
  for (int i = tid; i < N; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < nQ) {
      const float* v = reinterpret_cast<const float*>(geo + ((size_t)b * nQp + i) * 9 + 2);
      unsigned code = 0;
      unsigned qv[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float c = 0.5f * (v[6 + a] + v[12 + a]);
        const float ext = mx[a] - mn[a];
        float t = (ext > 0.f && c == c) ? (c - mn[a]) / ext : 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        qv[a] = (unsigned)(t * 1023.0f);
      }
#pragma unroll
      for (int bit = 0; bit < 10; ++bit)
#pragma unroll
        for (int a = 0; a < 3; ++a) code |= ((qv[a] >> bit) & 1u) << (3 * bit + a);
      key = ((unsigned long long)code << 32) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < N; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], c = keys[ixj];
          const bool asc = (i & k) == 0;
          if ((a > c) == asc) { keys[i] = c; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < nQ; i += blockDim.x) out[i] = (int)(keys[i] & 0xffffffffu);
}

// dense fp32 dbias [B][nQ][nK][4] -> scaled fp16 rows [(b*nQ + q)*4 + h][nKp]  (C-ABI helper only)
__global__ void rpe_dtables_dense_pack_kernel(const float4* __restrict__ ds4, size_t pairs, int nK, int nKp,
                                              __half* __restrict__ dsb, const unsigned* absmax_bits) {
  const float sc = scale_of(*absmax_bits, 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < pairs; i += (size_t)gridDim.x * blockDim.x) {
    const float4 d = __ldg(ds4 + i);
    const size_t bq = i / nK;
    const int k = (int)(i - bq * nK);
    __half* dst = dsb + bq * 4 * (size_t)nKp + k;
    dst[0] = __float2half_rn(d.x * sc);
    dst[(size_t)nKp] = __float2half_rn(d.y * sc);
    dst[2 * (size_t)nKp] = __float2half_rn(d.z * sc);
    dst[3 * (size_t)nKp] = __float2half_rn(d.w * sc);
  }
}

}  // namespace dt3

// developer aid: VDETR_DT_CLOCKS=1 accumulates per-phase cycle counts (read with vdetr_debug_dt_clocks)
static unsigned long long* g_dt_clocks = nullptr;
extern "C" int vdetr_debug_dt_clocks(unsigned long long* out8) {
  if (!g_dt_clocks) return VDETR_ERR_BAD_ARG;
  VDETR_CUDA_TRY(cudaMemcpy(out8, g_dt_clocks, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  VDETR_CUDA_TRY(cudaMemset(g_dt_clocks, 0, 8 * sizeof(unsigned long long)));
  return 0;
}

// scratch of the dTables pass: query order + one zero-padded private table per CTA
size_t rpe_dtables_scratch_bytes(const VdetrXattnShape* s) {
  const int P3 = s->grid_n + 2;
  return vdetr_align_up((size_t)s->B * s->nQ * sizeof(int), 1024) + 1024 +
         vdetr_align_up((size_t)vdetr_num_sms() * 8 * P3 * P3 * P3 * 4 * sizeof(float), 1024) + vdetr_align_up(rpe_dt6_priv_bytes(), 1024);
}

// dsb: scale * dS as fp16 rows [(b*nQp + q)*4 + h][nKp]; scale derives from *absmax_bits (vdetr_grad_scale, or the
// dense helper's own rule when dense_scale != 0).  dtables [8][n][n][n][4] is fully overwritten.
int rpe_dtables_launch(const VdetrXattnShape* s, int nQp, int nKp, const float4* xyz4, const float4* geo, const __half* dsb,
                       const unsigned* absmax_bits, int dense_scale, float* dtables, void* scratch, size_t scratch_bytes,
                       cudaStream_t st) {
  const int n = s->grid_n;
  if (s->B == 0 || s->nQ == 0 || s->nK == 0) {
    VDETR_CUDA_TRY(cudaMemsetAsync(dtables, 0, (size_t)8 * n * n * n * 4 * sizeof(float), st));
    return 0;
  }
  if (n < 1 || n > dt3::MAX_N) return VDETR_ERR_UNSUPPORTED;
  if (nKp % 4 != 0) return VDETR_ERR_BAD_ARG;                // dS rows are read 4 keys (8 bytes) at a time
  if (!scratch || scratch_bytes < rpe_dtables_scratch_bytes(s)) return VDETR_ERR_WORKSPACE;
  const size_t smem = dt3::smem_bytes(n);
  if (smem > 232448) return VDETR_ERR_UNSUPPORTED;
  uint8_t* w = reinterpret_cast<uint8_t*>(scratch);
  int* qperm = reinterpret_cast<int*>(w);
  int* slow_count = reinterpret_cast<int*>(w + vdetr_align_up((size_t)s->B * s->nQ * sizeof(int), 1024));
  float* priv = reinterpret_cast<float*>(w + vdetr_align_up((size_t)s->B * s->nQ * sizeof(int), 1024) + 1024);
  float* priv6 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(priv) +
                                          vdetr_align_up((size_t)vdetr_num_sms() * 8 * (n + 2) * (n + 2) * (n + 2) * 4 * sizeof(float), 1024));

  dt3::Params P = {};
  P.B = s->B; P.nQ = s->nQ; P.nK = s->nK; P.nQp = nQp; P.nKp = nKp; P.n = n; P.R = n + 1; P.P3 = n + 2;
  P.qblocks = (s->nQ + dt3::QB - 1) / dt3::QB;
  P.kchunks = (s->nK + dt3::KC - 1) / dt3::KC;
  P.units = s->B * P.qblocks * P.kchunks;
  P.dense_scale = dense_scale;
  P.log_scale = s->log_scale;
  P.c1 = (float)n / (2.0f * 3.0f * s->max_value);
  P.c0 = 0.5f * (float)(n - 1);
  P.xyz4 = xyz4; P.geo = geo; P.dsb = dsb; P.qperm = qperm; P.absmax_bits = absmax_bits; P.priv = priv;
  static const bool want_clocks = []() { const char* e = getenv("VDETR_DT_CLOCKS"); return e && e[0] == '1'; }();
  if (want_clocks && !g_dt_clocks) {
    VDETR_CUDA_TRY(cudaMalloc(&g_dt_clocks, 8 * sizeof(unsigned long long)));
    VDETR_CUDA_TRY(cudaMemset(g_dt_clocks, 0, 8 * sizeof(unsigned long long)));
  }
  P.phase_clocks = want_clocks ? g_dt_clocks : nullptr;
  static const int4 cost = []() {
    int4 k = make_int4(60, 200, 5, 600);
    if (const char* e = getenv("VDETR_DT_COST")) sscanf(e, "%d,%d,%d,%d", &k.x, &k.y, &k.z, &k.w);
    return k;
  }();
  P.cost = cost;
  P.slow_count = slow_count;
  // default (6): dt6, the dense tcgen05 contraction (rpe_dtables_umma.cu), for axis-aligned boxes + dt3 for the others (it
  // exits at once when there are none); VDETR_DT_IMPL=3: the dt3 kernel for every query; =5: dt5 (mma.sync accumulation) + dt3
  const char* impl_env = getenv("VDETR_DT_IMPL");          // read per call: tests switch it at run time
  const int impl = (impl_env && impl_env[0] == '5') ? 5 : (impl_env && impl_env[0] == '3') ? 3 : 6;
  const bool use_dt5 = impl == 5;
  P.only_slow = impl != 3 ? 1 : 0;
  const int grid3 = P.units < vdetr_num_sms() ? P.units : vdetr_num_sms();
  dt3::Params P4 = P;
  P4.qblocks = (s->nQ + dt5::QB - 1) / dt5::QB;
  P4.kchunks = (s->nK + dt5::KC - 1) / dt5::KC;
  P4.units = s->B * P4.qblocks * P4.kchunks;
  const int grid4 = P4.units < vdetr_num_sms() ? P4.units : vdetr_num_sms();
  const size_t smem4 = dt5::smem_bytes(n);
  if (use_dt5 && smem4 > 232448) return VDETR_ERR_UNSUPPORTED;
  const int copies = (use_dt5 && grid4 > grid3) ? grid4 : grid3;
  const size_t copy_bytes = (size_t)8 * P.P3 * P.P3 * P.P3 * 4 * sizeof(float);

  VdetrTimingScope timing(VDETR_T_DTABLES, st);
  VDETR_CUDA_TRY(cudaMemsetAsync(priv, 0, copy_bytes * copies, st));
  VDETR_CUDA_TRY(cudaMemsetAsync(slow_count, 0, sizeof(int), st));
  dt3::rpe_dtables_qorder_kernel<<<s->B, 1024, 0, st>>>(geo, s->nQ, nQp, qperm, slow_count);
  VDETR_LAUNCH_CHECK();
  if (impl == 5) {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(dt5::rpe_dtables_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
    dt5::rpe_dtables_win_kernel<<<grid4, dt5::THREADS, smem4, st>>>(P4);
    VDETR_LAUNCH_CHECK();
  }
  VDETR_CUDA_TRY(cudaFuncSetAttribute(dt3::rpe_dtables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dt3::rpe_dtables_kernel<<<grid3, dt3::THREADS, smem, st>>>(P);
  VDETR_LAUNCH_CHECK();
  const int total = 8 * n * n * n * 4;
  dt3::rpe_dtables_reduce_kernel<<<(total + 31) / 32, 256, 0, st>>>(priv, copies, n, P.P3, dtables, absmax_bits, dense_scale);
  VDETR_LAUNCH_CHECK();
  if (impl == 6) return rpe_dt6_launch(s, nQp, nKp, xyz4, geo, dsb, absmax_bits, dense_scale, dtables, 1, priv6, st);
  return 0;
}

// C-ABI helper behind vdetr_rpe_dtables: dense dS [B,nQ,nK,4] -> dTables (packs xyz / geometry / fp16 dS itself).
size_t rpe_dtables_workspace(const VdetrXattnShape* s) {
  const size_t nKp = vdetr_align_up((size_t)s->nK, 64);
  return vdetr_align_up((size_t)s->B * nKp * 16, 1024) + vdetr_align_up((size_t)s->B * s->nQ * 9 * 16, 1024) +
         vdetr_align_up((size_t)s->B * s->nQ * 4 * nKp * 2, 1024) + 1024 + rpe_dtables_scratch_bytes(s);
}
int rpe_dtables_dense(const VdetrXattnShape* s, const float* xyz, const float* ref, const float* ang, const float* ds4,
                      float* dtables, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < rpe_dtables_workspace(s)) return VDETR_ERR_WORKSPACE;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  const int nKp = (int)vdetr_align_up((size_t)s->nK, 64);
  VdetrPack pk = {};
  pk.B = s->B; pk.nQ = s->nQ; pk.nK = s->nK; pk.nQp = s->nQ; pk.nKp = nKp; pk.kvh = 1; pk.has_bias = 1;
  pk.xyz = xyz; pk.ref = ref; pk.ang = s->rotate ? ang : nullptr;
  size_t o = 0;
  pk.xyz4 = reinterpret_cast<float4*>(w + o); o += vdetr_align_up((size_t)s->B * nKp * 16, 1024);
  pk.geo = reinterpret_cast<float4*>(w + o); o += vdetr_align_up((size_t)s->B * s->nQ * 9 * 16, 1024);
  __half* dsb = reinterpret_cast<__half*>(w + o);
  const size_t dsb_bytes = (size_t)s->B * s->nQ * 4 * nKp * 2;
  o += vdetr_align_up(dsb_bytes, 1024);
  unsigned* absmax = reinterpret_cast<unsigned*>(w + o); o += 1024;
  if (s->B == 0 || s->nQ == 0 || s->nK == 0)
    return rpe_dtables_launch(s, s->nQ, nKp, nullptr, nullptr, nullptr, nullptr, 1, dtables, w + o, ws_bytes - o, st);
  vdetr_pack_kernel<<<vdetr_num_sms(), 256, 0, st>>>(pk);
  VDETR_LAUNCH_CHECK();
  const size_t pairs = (size_t)s->B * s->nQ * s->nK;
  VDETR_CUDA_TRY(cudaMemsetAsync(absmax, 0, 4, st));
  VDETR_CUDA_TRY(cudaMemsetAsync(dsb, 0, dsb_bytes, st));          // padding columns
  vdetr_absmax_kernel<<<vdetr_num_sms() * 2, 256, 0, st>>>(ds4, pairs * 4, absmax);
  VDETR_LAUNCH_CHECK();
  dt3::rpe_dtables_dense_pack_kernel<<<vdetr_num_sms() * 4, 256, 0, st>>>(reinterpret_cast<const float4*>(ds4), pairs, s->nK, nKp, dsb, absmax);
  VDETR_LAUNCH_CHECK();
  return rpe_dtables_launch(s, s->nQ, nKp, pk.xyz4, pk.geo, dsb, absmax, 1, dtables, w + o, ws_bytes - o, st);
}
