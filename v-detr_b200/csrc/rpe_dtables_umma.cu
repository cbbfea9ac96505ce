// dt6: dTables of AXIS-ALIGNED boxes as a dense tcgen05 contraction (the default dTables kernel; boxes that are not axis
// aligned go through dt3 with only_slow = 1, rpe_dtables.cu).
//
// For an axis-aligned box the trilinear weight of vertex (sx, sy, sz) in {+,-}^3 factors per axis,
//     w(q, k; z, y, x) = hz^{sz}(q,k)[z] * hy^{sy}(q,k)[y] * hx^{sx}(q,k)[x],
// every h a "hat" with (at most) two non-zero entries among the n <= 10 table points of that axis, so
//     dT_{sx,sy,sz}[z][y][x][h] = sum_pairs ( hz^{sz}[z] hy^{sy}[y] ) * ( hx^{sx}[x] dS[h] )
// is a GEMM over the pair index:  D_{sz,sy} [100 (z,y) x 80 (sx, x, h)]  +=  A_{sz,sy} [100 x pairs] * B [pairs x 80].
// The operands are 96 % zeros -- but scattering the 16 + 16 non-zero fp16 values of a pair into shared-memory tiles costs
// ~20 instructions per vertex evaluation, where sorting the pairs by table cell and accumulating them in registers (dt3)
// costs ~300, and the tensor pipe is otherwise idle in this phase of the step.  The four accumulators (4 x 80 fp32 TMEM
// columns x 128 lanes) stay in tensor memory for the whole launch: no atomics, no flushes; every CTA leaves one private
// copy and a second kernel sums the copies in a fixed order (bit-reproducible).
//
//   stage   = 64 pairs (one query x 64 consecutive keys): four A tiles [104 rows x 64 pairs] + one B tile [88 x 64], fp16,
//             K-major rows of 128 B with the 128-byte swizzle (the layout TMA would write).  Rows >= 100 of A / >= 80 of B
//             are dump rows for corners outside the table (zero padding of grid_sample), never read as results: the MMA
//             (M = 128) reads 24 rows past each A tile, which only feeds accumulator lanes 104..127 that nobody reads.
//   warps   : 3 stages x 4 producer warps (lane = pair; a warp owns 32 pairs and the z sign `vh`: it writes A_{vh,+},
//             A_{vh,-} and the x^{vh} half of B), four MMA warps (one lane each issues the 4 K-steps of its variant per stage:
//             tcgen05.mma M = 128, N = 80, K = 16).  A producer computes the next item while the MMAs of its stage run,
//             then zeroes the 16 entries it wrote last time and writes the new ones.
//   bound   : 40 cycles per MMA (128 x 80 x 16 MACs) x 16 = 640 cycles per 64 pairs and SM.
#include "rpe_internal.h"
#include "rpe_fast.cuh"
#include <stdlib.h>

namespace dt6 {

using namespace tc;

constexpr int STAGES = 3, KS = 64;
constexpr int TP = 10;                                  // table points per axis the tiles are laid out for (n <= TP)
constexpr int A_ROWS = 104, A_BYTES = A_ROWS * 128;     // rows 100..103: dump
constexpr int B_ROWS = 88, B_BYTES = B_ROWS * 128;      // rows 80..87: dump
constexpr int NCOL = 2 * TP * 4;                        // 80 accumulator columns: (x sign, x point, head)
constexpr int STAGE_BYTES = 4 * A_BYTES + B_BYTES;      // 64512
constexpr int PROD_WARPS = STAGES * 4;
constexpr int MMA_WARPS = 4;                            // one per (z sign, y sign) variant: issuing an MMA costs ~100 cycles
constexpr int THREADS = (PROD_WARPS + MMA_WARPS) * 32;
constexpr int COPY_FLOATS = 4 * TP * TP * NCOL;         // one private copy: [variant][z * 10 + y][80]
static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "tiles start on 1024-byte swizzle atoms");
static_assert(3 * 1024 <= B_BYTES, "the MMA over-read of the last A tile stays inside the stage");

struct Params {
  int B, nQ, nK, nQp, nKp, n, KT;
  long long items;                                      // B * nQ * KT stages of 64 pairs
  float log_scale, c1, c0;
  const float4* xyz4;                                   // [B][nKp]
  const float4* geo;                                    // [B][nQp][9]
  const __half* dsb;                                    // [(b*nQp + q)*4 + h][nKp]   scale * dS
  float* priv;                                          // [gridDim.x][COPY_FLOATS]
  unsigned long long* clk;                              // developer: [8] phase cycle sums (VDETR_DT_CLOCKS=1), else null
  int dbg;                                              // developer: 1 = no MMAs, 2 = no producer work, 4 = no loads
};

__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
// byte offset of (row, this lane's pair column) inside a K-major 128-byte-swizzled tile; lanec = (chunk << 4) | (elem << 1)
__device__ __forceinline__ uint32_t sw_off(uint32_t row, uint32_t lanec) { return ((row << 7) | ((row & 7u) << 4)) ^ lanec; }

// one axis: base table point n0 (points n0 and n0 + 1 carry 1 - f and f) -- the forward's arithmetic (rpe_axis_fast)
__device__ __forceinline__ void axis_pt(float d, float ls, float c1, float c0, int n, int& n0, float& f) {
  const float t = lg2_approx(fmaf(fabsf(d), ls, 1.0f)) * c1;
  float ts = copysignf(t, d);
  ts = fminf(fmaxf(ts, -c0 - 1.5f), (float)n - c0 + 0.5f);
  const float p = ts + c0;
  const float r = (p - 0.5f) + rpe::MAGIC;
  f = p - (r - rpe::MAGIC);
  n0 = __float_as_int(r) - rpe::MAGIC_BITS;
}

struct Raw {                    // what a producer lane loads for one pair
  float4 kx;                    // key xyz
  float zq, yp, ym, xq;         // the query's z^{vh}, y+, y-, x^{vh} box faces
  int fast;                     // axis-aligned box
  unsigned short d[4];          // scaled fp16 dS of the 4 heads
};

__global__ void __launch_bounds__(THREADS, 1) rpe_dtables_umma_kernel(const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bar_full = bars;                 // [STAGES] the producers of a stage have written it
  uint64_t* bar_empty = bars + STAGES;       // [STAGES] the MMAs reading a stage are done
  uint64_t* bar_done = bars + 2 * STAGES;    // all MMAs of the CTA are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long tk0 = P.clk ? clock64() : 0;

  {                                          // tiles start as zeros
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < STAGES * STAGE_BYTES / 16; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == PROD_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s, 128); mbar_init(bar_empty + s, MMA_WARPS); }
      mbar_init(bar_done, MMA_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long tk1 = P.clk ? clock64() : 0;

  // this CTA's contiguous share of the items (all items cost the same)
  const long long i_begin = P.items * blockIdx.x / gridDim.x, i_end = P.items * (blockIdx.x + 1) / gridDim.x;
  const int my_items = (int)(i_end - i_begin);

  if (warp < PROD_WARPS) {
    // ------------------------------------------------------------------------------------------ producers
    const int stage = warp >> 2, pw = warp & 3, khalf = pw & 1, vh = pw >> 1;
    const uint32_t kloc = (uint32_t)(khalf * 32 + lane);
    const uint32_t lanec = ((kloc >> 3) << 4) | ((kloc & 7u) << 1);
    const uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES);
    const uint32_t tA0 = sbase + (uint32_t)(vh * 2) * A_BYTES, tA1 = tA0 + A_BYTES, tB = sbase + 4 * A_BYTES;
    const int n = P.n;
    uint32_t clr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) clr[i] = tA0 + sw_off(100, lanec);

    // decode the first item of this stage
    long long g0 = i_begin + stage;
    int kt = (int)(g0 % P.KT);
    long long r = g0 / P.KT;
    int q = (int)(r % P.nQ), b = (int)(r / P.nQ);

    auto load = [&](Raw& w, int b_, int q_, int kt_) {
      const int k = kt_ * KS + (int)kloc;
      const float* g = reinterpret_cast<const float*>(P.geo + ((size_t)b_ * P.nQp + q_) * 9);
      w.zq = __ldg(g + vh * 4 + 2); w.yp = __ldg(g + 1); w.ym = __ldg(g + 5); w.xq = __ldg(g + vh * 4);
      w.fast = __float_as_int(__ldg(g + 3));
      w.kx = __ldg(P.xyz4 + (size_t)b_ * P.nKp + k);
      const unsigned short* dp = reinterpret_cast<const unsigned short*>(P.dsb) + ((size_t)b_ * P.nQp + q_) * 4 * P.nKp + k;
#pragma unroll
      for (int h = 0; h < 4; ++h) w.d[h] = __ldg(dp + (size_t)h * P.nKp);
    };
    // Two register sets in ping-pong: the set an item was computed from is reloaded (after the stage has been handed to the
    // MMA warp) with the item two rounds ahead, so a load has a whole round (~2000 cycles) to land, no register is copied
    // while its load is in flight, and the MEMBAR inside fence.proxy.async never waits for a young load.
    auto advance = [&]() {
      kt += STAGES;
      while (kt >= P.KT) { kt -= P.KT; if (++q == P.nQ) { q = 0; ++b; } }
    };
    Raw r0, r1;
    int kt0 = kt, kt1 = 0;
    if (stage < my_items) load(r0, b, q, kt);
    advance();
    if (stage + STAGES < my_items) { load(r1, b, q, kt); kt1 = kt; }

    long long c_comp = 0, c_wait = 0, c_store = 0, c_load = 0;
    auto step = [&](Raw& cur, int& ckt, int it) {
      const long long t0 = P.clk ? clock64() : 0;
      const bool active = (ckt * KS + (int)kloc) < P.nK && cur.fast != 0;   // inside nK, axis-aligned box
      const float act = active ? 1.f : 0.f;
      int nz, nyp, nym, nx;
      float fz, fyp, fym, fx;
      axis_pt(cur.zq - cur.kx.z, P.log_scale, P.c1, P.c0, n, nz, fz);
      axis_pt(cur.yp - cur.kx.y, P.log_scale, P.c1, P.c0, n, nyp, fyp);
      axis_pt(cur.ym - cur.kx.y, P.log_scale, P.c1, P.c0, n, nym, fym);
      axis_pt(cur.xq - cur.kx.x, P.log_scale, P.c1, P.c0, n, nx, fx);

      uint32_t adr[16], val[16];
      const float wz0 = (1.f - fz) * act, wz1 = fz * act;
      const bool vz0 = (unsigned)nz < (unsigned)n, vz1 = (unsigned)(nz + 1) < (unsigned)n;
#pragma unroll
      for (int ys = 0; ys < 2; ++ys) {
        const int ny = ys ? nym : nyp;
        const float fy = ys ? fym : fyp;
        const bool vy0 = (unsigned)ny < (unsigned)n, vy1 = (unsigned)(ny + 1) < (unsigned)n;
        const uint32_t tile = ys ? tA1 : tA0;
        const int r00 = nz * TP + ny;
        const uint32_t p0 = pack_f16x2(wz0 * (1.f - fy), wz0 * fy), p1 = pack_f16x2(wz1 * (1.f - fy), wz1 * fy);
        adr[ys * 4 + 0] = tile + sw_off((vz0 && vy0) ? (uint32_t)r00 : 100u, lanec);
        adr[ys * 4 + 1] = tile + sw_off((vz0 && vy1) ? (uint32_t)(r00 + 1) : 101u, lanec);
        adr[ys * 4 + 2] = tile + sw_off((vz1 && vy0) ? (uint32_t)(r00 + TP) : 102u, lanec);
        adr[ys * 4 + 3] = tile + sw_off((vz1 && vy1) ? (uint32_t)(r00 + TP + 1) : 103u, lanec);
        val[ys * 4 + 0] = p0; val[ys * 4 + 1] = p0 >> 16; val[ys * 4 + 2] = p1; val[ys * 4 + 3] = p1 >> 16;
      }
      {
        // hx[x] * dS[h]: fp32 products rounded once to fp16 (dS arrives as scaled fp16)
        const float2 d01 = __half22float2(__halves2half2(__ushort_as_half(cur.d[0]), __ushort_as_half(cur.d[1])));
        const float2 d23 = __half22float2(__halves2half2(__ushort_as_half(cur.d[2]), __ushort_as_half(cur.d[3])));
        const float w0 = (1.f - fx) * act, w1 = fx * act;
        const bool vx0 = (unsigned)nx < (unsigned)n, vx1 = (unsigned)(nx + 1) < (unsigned)n;
        const uint32_t c0 = vx0 ? (uint32_t)((vh * TP + nx) * 4) : 80u, c1 = vx1 ? (uint32_t)((vh * TP + nx + 1) * 4) : 84u;
        const uint32_t ua01 = pack_f16x2(w0 * d01.x, w0 * d01.y), ua23 = pack_f16x2(w0 * d23.x, w0 * d23.y);
        const uint32_t ub01 = pack_f16x2(w1 * d01.x, w1 * d01.y), ub23 = pack_f16x2(w1 * d23.x, w1 * d23.y);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          adr[8 + h] = tB + sw_off(c0 + h, lanec);
          adr[12 + h] = tB + sw_off(c1 + h, lanec);
        }
        val[8] = ua01; val[9] = ua01 >> 16; val[10] = ua23; val[11] = ua23 >> 16;
        val[12] = ub01; val[13] = ub01 >> 16; val[14] = ub23; val[15] = ub23 >> 16;
      }

      const int round = it / STAGES;
      // (the stores depend on everything computed above, so t1 is taken after the compute phase has retired)
      const long long t1 = P.clk ? clock64() + (long long)(__float_as_int(fz + fx + fyp + fym) & 0) + (adr[0] & 0) + (adr[15] & 0) + (val[15] & 0) : 0;
      if (round > 0) mbar_wait(bar_empty + stage, (round - 1) & 1);        // the MMAs that read this stage are done
      const long long t2 = P.clk ? clock64() : 0;
      if (!(P.dbg & 2)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sts16(clr[i], 0u);
#pragma unroll
        for (int i = 0; i < 16; ++i) { sts16(adr[i], val[i]); clr[i] = adr[i]; }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_full + stage);
      const long long t3 = P.clk ? clock64() : 0;
      // reload this register set with the item two rounds ahead
      advance();
      if (it + 2 * STAGES < my_items && !(P.dbg & 4)) { load(cur, b, q, kt); ckt = kt; }
      if (P.clk) { const long long t4 = clock64(); c_comp += t1 - t0; c_wait += t2 - t1; c_store += t3 - t2; c_load += t4 - t3; }
    };
    for (int it = stage; it < my_items; it += 2 * STAGES) {
      step(r0, kt0, it);
      if (it + STAGES < my_items) step(r1, kt1, it + STAGES);
    }
    if (P.clk && tid == 0) { atomicAdd(P.clk + 6, (unsigned long long)(tk1 - tk0)); atomicAdd(P.clk + 7, (unsigned long long)(clock64() - tk1)); }
    if (P.clk && lane == 0) {
      atomicAdd(P.clk + 0, (unsigned long long)c_comp); atomicAdd(P.clk + 1, (unsigned long long)c_wait);
      atomicAdd(P.clk + 2, (unsigned long long)c_store); atomicAdd(P.clk + 3, (unsigned long long)c_load);
    }
  } else {
    // ------------------------------------------------------------------------------------------ MMA issuers
    // Warp PROD_WARPS + v issues the 4 K-steps of variant v of every stage.  (One lane issuing all 16 MMAs of a stage was
    // the bottleneck of the first version: ~95 cycles of descriptor moves and issue per MMA against 40 cycles of execution.)
    if (lane == 0) {
      const int v = warp - PROD_WARPS;
      const uint32_t idesc = umma_idesc_f16((P.dbg & 16) ? 64 : 128, (P.dbg & 8) ? 48 : ((P.dbg & 32) ? 160 : NCOL));   // (developer timing experiments)
      const uint32_t td = tmem_base + v * NCOL;
      uint64_t da[STAGES], db[STAGES];
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const uint32_t sbase = smem_u32(smem + s * STAGE_BYTES);
        da[s] = umma_desc_sw128(sbase + v * A_BYTES);
        db[s] = umma_desc_sw128(sbase + 4 * A_BYTES);
      }
      long long m_wait = 0, m_issue = 0;
      for (int it0 = 0; it0 < my_items; it0 += STAGES) {
        const uint32_t par = (uint32_t)(it0 / STAGES) & 1u;
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
          if (it0 + s < my_items) {
            const long long t0 = P.clk ? clock64() : 0;
            mbar_wait(bar_full + s, par);
            const long long t1 = P.clk ? clock64() : 0;
            tc_fence_after();
            if (!(P.dbg & 1)) {
#pragma unroll
              for (int kk = 0; kk < KS / 16; ++kk)
                umma_bf16(td, da[s] + (uint64_t)(kk * 2), db[s] + (uint64_t)(kk * 2), idesc, (it0 + s) > 0 || kk > 0);
            }
            umma_commit(bar_empty + s);
            if (P.clk) { m_wait += t1 - t0; m_issue += clock64() - t1; }
          }
        }
      }
      umma_commit(bar_done);
      if (P.clk) { atomicAdd(P.clk + 4, (unsigned long long)m_wait); atomicAdd(P.clk + 5, (unsigned long long)m_issue); }
    }
    __syncwarp();
  }

  // -------------------------------------------------------------------------------------------- epilogue: TMEM -> private copy
  if (warp < 8) {
    mbar_wait_relaxed(bar_done, 0);
    tc_fence_after();
    const int quad = warp & 3, row = quad * 32 + lane;
    float* dst = P.priv + (size_t)blockIdx.x * COPY_FLOATS;
#pragma unroll 1
    for (int vv = 0; vv < 2; ++vv) {
      const int v = (warp >> 2) * 2 + vv;
#pragma unroll 1
      for (int c = 0; c < NCOL / 16; ++c) {
        uint32_t rr[16];
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(v * NCOL + c * 16), rr);
        tmem_ld_wait();
        if (row < TP * TP) {
          float4* o = reinterpret_cast<float4*>(dst + ((size_t)v * TP * TP + row) * NCOL + c * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PROD_WARPS) tmem_dealloc<512>(tmem_base);
}

// out[vertex][z][y][x][h] (+)= (1 / scale) * sum over the private copies, in a fixed order.  A CTA owns 32 consecutive
// outputs; its 8 warps split the copies.
__global__ void __launch_bounds__(256) rpe_dtables_umma_reduce_kernel(const float* __restrict__ priv, int copies, int n,
                                                                      float* __restrict__ out, const unsigned* absmax_bits, int dense,
                                                                      int accumulate) {
  __shared__ float part[8][32];
  const float inv = 1.0f / vdetr_dt_scale(*absmax_bits, dense);
  const int total = 8 * n * n * n * 4;
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
      // vertex sign table: 0:(+,+,-) 1:(+,-,-) 2:(-,-,-) 3:(-,+,-) 4..7: the same with z +
      const int xs = ((v & 3) >= 2) ? 1 : 0, ys = ((v & 3) == 1 || (v & 3) == 2) ? 1 : 0, zs = (v < 4) ? 1 : 0;
      const size_t src = ((size_t)(zs * 2 + ys) * TP * TP + z * TP + y) * NCOL + (xs * TP + x) * 4 + h;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int c = warp;
      for (; c + 24 < copies; c += 32) {
        s0 += priv[(size_t)c * COPY_FLOATS + src]; s1 += priv[(size_t)(c + 8) * COPY_FLOATS + src];
        s2 += priv[(size_t)(c + 16) * COPY_FLOATS + src]; s3 += priv[(size_t)(c + 24) * COPY_FLOATS + src];
      }
      for (; c < copies; c += 8) s0 += priv[(size_t)c * COPY_FLOATS + src];
      s = (s0 + s1) + (s2 + s3);
    }
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][lane];
      out[i] = (accumulate ? out[i] : 0.f) + t * inv;
    }
    __syncthreads();
  }
}

}  // namespace dt6

static unsigned long long* g_dt6_clocks = nullptr;
extern "C" int vdetr_debug_dt6_clocks(unsigned long long* out8) {
  if (!g_dt6_clocks) return VDETR_ERR_BAD_ARG;
  VDETR_CUDA_TRY(cudaMemcpy(out8, g_dt6_clocks, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  VDETR_CUDA_TRY(cudaMemset(g_dt6_clocks, 0, 8 * sizeof(unsigned long long)));
  return 0;
}

size_t rpe_dt6_priv_bytes() { return (size_t)vdetr_num_sms() * dt6::COPY_FLOATS * sizeof(float); }

// Accumulates the axis-aligned queries of the call and adds the result to `dtables` (accumulate != 0) or overwrites it.
int rpe_dt6_launch(const VdetrXattnShape* s, int nQp, int nKp, const float4* xyz4, const float4* geo, const __half* dsb,
                   const unsigned* absmax_bits, int dense_scale, float* dtables, int accumulate, float* priv, cudaStream_t st) {
  const int n = s->grid_n;
  if (n < 1 || n > dt6::TP || nKp % dt6::KS != 0) return VDETR_ERR_UNSUPPORTED;
  dt6::Params P = {};
  P.B = s->B; P.nQ = s->nQ; P.nK = s->nK; P.nQp = nQp; P.nKp = nKp; P.n = n;
  P.KT = (s->nK + dt6::KS - 1) / dt6::KS;
  P.items = (long long)s->B * s->nQ * P.KT;
  P.log_scale = s->log_scale;
  P.c1 = (float)n / (2.0f * 3.0f * s->max_value);
  P.c0 = 0.5f * (float)(n - 1);
  P.xyz4 = xyz4; P.geo = geo; P.dsb = dsb; P.priv = priv;
  { const char* e = getenv("VDETR_DT6_DBG"); P.dbg = e ? atoi(e) : 0; }
  static const bool want_clocks = []() { const char* e = getenv("VDETR_DT_CLOCKS"); return e && e[0] == '1'; }();
  if (want_clocks && !g_dt6_clocks) {
    VDETR_CUDA_TRY(cudaMalloc(&g_dt6_clocks, 8 * sizeof(unsigned long long)));
    VDETR_CUDA_TRY(cudaMemset(g_dt6_clocks, 0, 8 * sizeof(unsigned long long)));
  }
  P.clk = want_clocks ? g_dt6_clocks : nullptr;
  const int grid = (int)(P.items < (long long)vdetr_num_sms() ? P.items : (long long)vdetr_num_sms());
  const size_t smem = (size_t)dt6::STAGES * dt6::STAGE_BYTES + 256 + 1024;
  VDETR_CUDA_TRY(cudaFuncSetAttribute(dt6::rpe_dtables_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dt6::rpe_dtables_umma_kernel<<<grid, dt6::THREADS, smem, st>>>(P);
  VDETR_LAUNCH_CHECK();
  const int total = 8 * n * n * n * 4;
  dt6::rpe_dtables_umma_reduce_kernel<<<(total + 31) / 32, 256, 0, st>>>(priv, grid, n, dtables, absmax_bits, dense_scale, accumulate);
  VDETR_LAUNCH_CHECK();
  return 0;
}
