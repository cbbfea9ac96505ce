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
//   stage   = 64 pairs (one query x 64 consecutive keys): four A tiles [104 rows x 64 pairs], fp16, K-major rows of 128 B with
//             the 128-byte swizzle (the layout TMA would write), and one B tile stored MN-MAJOR ([64 pairs][128 columns] in two
//             64-column blocks, 128-byte swizzle): the 4 heads of one x table point are 8 contiguous bytes of a pair's row, so
//             a producer writes them with ONE 64-bit store.  Rows >= 100 of A / columns >= 80 of B are dump slots for corners
//             outside the table (zero padding of grid_sample), never read as results: the MMA (M = 128) reads 24 rows past
//             each A tile, which only feeds accumulator lanes 104..127 that nobody reads.
//   warps   : 3 stages x 6 producer warps (lane = pair; a warp owns 32 pairs and one role: the A tiles of y+, the A tiles of
//             y-, or the B tile -- see produce<>), four MMA warps (one lane each issues the 4 K-steps of its variant per
//             stage: tcgen05.mma M = 128, N = 80, K = 16).  A producer computes the next item while the MMAs of its stage
//             run, then zeroes the entries it wrote last time and writes the new ones.  The producers are bound by
//             instruction issue, so the roles are cut to minimise instructions per pair (6 + 2 axis evaluations, 16 + 4
//             stores); the first version (4 warps per stage, 8 + 8 axes, 32 16-bit stores per pair) kept the MMA warps
//             waiting 40 % of the time.
//   bound   : 40 cycles per MMA (128 x 80 x 16 MACs) x 16 = 640 cycles per 64 pairs and SM.
#include "rpe_internal.h"
#include "rpe_fast.cuh"
#include <stdlib.h>

namespace dt6 {

using namespace tc;

constexpr int STAGES = 3, KS = 64;
constexpr int TP = 10;                                  // table points per axis the tiles are laid out for (n <= TP)
constexpr int A_ROWS = 104, A_BYTES = A_ROWS * 128;     // rows 100..103: dump
constexpr int B_BYTES = 2 * KS * 128;                   // MN-major: two 64-column blocks of [64 pairs][128 B]; columns 80..87: dump
constexpr int NCOL = 2 * TP * 4;                        // 80 accumulator columns: (x sign, x point, head)
constexpr int STAGE_BYTES = 4 * A_BYTES + B_BYTES;      // 64512
constexpr int WPS = 6;                                  // producer warps per stage
constexpr int PROD_WARPS = STAGES * WPS;
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

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t lo, uint32_t hi) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}

// One producer warp: 32 pairs (lane = pair) of every STAGES-th item of the CTA, for one role:
//   ROLE 0 / 1 : the y sign ys = ROLE; writes the 4 entries of A_{z+,ys} and of A_{z-,ys}   (3 axes: z+, z-, y^{ys})
//   ROLE 2     : the B tile: for x+ and x-, the two x table points times the 4 heads of dS = four 64-bit stores (2 axes)
// Per step: compute the entries of the next item from registers loaded two rounds ago, wait until the MMAs that read the
// stage are done, zero the entries written last time, write the new ones, hand the stage over, reload the register set.
template <int ROLE>
__device__ __forceinline__ void produce(const Params& P, uint8_t* smem, uint64_t* bar_full, uint64_t* bar_empty, int stage, int khalf,
                                        int lane, long long i_begin, int my_items, long long tk0, long long tk1) {
  constexpr int NA = (ROLE == 2) ? 4 : 8;               // stores per step (16-bit for A, 64-bit for B)
  const uint32_t kloc = (uint32_t)(khalf * 32 + lane);
  const uint32_t lanec = ((kloc >> 3) << 4) | ((kloc & 7u) << 1);
  const uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES);
  const uint32_t tA0 = sbase + (uint32_t)ROLE * A_BYTES, tA1 = tA0 + 2 * A_BYTES;          // (z+, ys), (z-, ys)
  // MN-major B: this pair's 128-byte row of column block 0; the 4 heads of column group c (a multiple of 4) are 8 bytes
  const uint32_t brow = sbase + 4 * A_BYTES + kloc * 128u, bsw = kloc & 7u;
  auto b_off = [&](uint32_t c) { return brow + (c >> 6) * (uint32_t)(KS * 128) + ((((c & 63u) >> 3) ^ bsw) << 4) + ((c & 4u) << 1); };
  const int n = P.n;

  struct Raw {                    // what a lane loads for one pair
    float4 kx;                    // key xyz
    float f0, f1, f2;             // box faces: ROLE 0/1: z+, z-, y^{ys};  ROLE 2: x+, x-
    int fast;                     // axis-aligned box
    unsigned short d[4];          // ROLE 2: scaled fp16 dS of the 4 heads
  };
  // decode the first item of this stage
  long long g0 = i_begin + stage;
  int kt = (int)(g0 % P.KT);
  long long r = g0 / P.KT;
  int q = (int)(r % P.nQ), b = (int)(r / P.nQ);

  auto load = [&](Raw& w, int b_, int q_, int kt_) {
    const int k = kt_ * KS + (int)kloc;
    const float* g = reinterpret_cast<const float*>(P.geo + ((size_t)b_ * P.nQp + q_) * 9);
    if (ROLE < 2) { w.f0 = __ldg(g + 2); w.f1 = __ldg(g + 6); w.f2 = __ldg(g + ROLE * 4 + 1); }
    else { w.f0 = __ldg(g); w.f1 = __ldg(g + 4); }
    w.fast = __float_as_int(__ldg(g + 3));
    w.kx = __ldg(P.xyz4 + (size_t)b_ * P.nKp + k);
    if (ROLE == 2) {
      const unsigned short* dp = reinterpret_cast<const unsigned short*>(P.dsb) + ((size_t)b_ * P.nQp + q_) * 4 * P.nKp + k;
#pragma unroll
      for (int h = 0; h < 4; ++h) w.d[h] = __ldg(dp + (size_t)h * P.nKp);
    }
  };
  // Two register sets in ping-pong: the set an item was computed from is reloaded (after the stage has been handed to the
  // MMA warp) with the item two rounds ahead, so a load has a whole round to land, no register is copied while its load
  // is in flight, and the MEMBAR inside fence.proxy.async never waits for a young load.  The addresses written by one
  // step are the ones the next step clears: two address sets in ping-pong as well.
  auto advance = [&]() {
    kt += STAGES;
    while (kt >= P.KT) { kt -= P.KT; if (++q == P.nQ) { q = 0; ++b; } }
  };
  Raw r0, r1;
  int kt0 = kt, kt1 = 0;
  if (stage < my_items) load(r0, b, q, kt);
  advance();
  if (stage + STAGES < my_items) { load(r1, b, q, kt); kt1 = kt; }
  uint32_t adr0[NA], adr1[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) adr0[i] = adr1[i] = (ROLE == 2) ? b_off(80u) : tA0 + sw_off(100, lanec);

  long long c_comp = 0, c_wait = 0, c_store = 0, c_load = 0;
  auto step = [&](Raw& cur, int& ckt, int it, uint32_t (&adr)[NA], const uint32_t (&old)[NA]) {
    const long long t0 = P.clk ? clock64() : 0;
    const bool active = (ckt * KS + (int)kloc) < P.nK && cur.fast != 0;   // inside nK, axis-aligned box
    const float act = active ? 1.f : 0.f;
    uint32_t val[NA], hi2[4];                           // hi2: heads 2, 3 of the 64-bit B stores
    if (ROLE < 2) {
      int nz[2], ny;
      float fz[2], fy;
      axis_pt(cur.f0 - cur.kx.z, P.log_scale, P.c1, P.c0, n, nz[0], fz[0]);
      axis_pt(cur.f1 - cur.kx.z, P.log_scale, P.c1, P.c0, n, nz[1], fz[1]);
      axis_pt(cur.f2 - cur.kx.y, P.log_scale, P.c1, P.c0, n, ny, fy);
      const bool vy0 = (unsigned)ny < (unsigned)n, vy1 = (unsigned)(ny + 1) < (unsigned)n;
      const float wy0 = (1.f - fy) * act, wy1 = fy * act;
#pragma unroll
      for (int zs = 0; zs < 2; ++zs) {
        const uint32_t tile = zs ? tA1 : tA0;
        const bool vz0 = (unsigned)nz[zs] < (unsigned)n, vz1 = (unsigned)(nz[zs] + 1) < (unsigned)n;
        const float wz1 = fz[zs], wz0 = 1.f - wz1;
        const int r00 = nz[zs] * TP + ny;
        const uint32_t p0 = pack_f16x2(wz0 * wy0, wz0 * wy1), p1 = pack_f16x2(wz1 * wy0, wz1 * wy1);
        adr[zs * 4 + 0] = tile + sw_off((vz0 && vy0) ? (uint32_t)r00 : 100u, lanec);
        adr[zs * 4 + 1] = tile + sw_off((vz0 && vy1) ? (uint32_t)(r00 + 1) : 101u, lanec);
        adr[zs * 4 + 2] = tile + sw_off((vz1 && vy0) ? (uint32_t)(r00 + TP) : 102u, lanec);
        adr[zs * 4 + 3] = tile + sw_off((vz1 && vy1) ? (uint32_t)(r00 + TP + 1) : 103u, lanec);
        val[zs * 4 + 0] = p0; val[zs * 4 + 1] = p0 >> 16; val[zs * 4 + 2] = p1; val[zs * 4 + 3] = p1 >> 16;
      }
    } else {
      // hx[x] * dS[h]: fp32 products rounded once to fp16 (dS arrives as scaled fp16)
      const float2 d01 = __half22float2(__halves2half2(__ushort_as_half(cur.d[0]), __ushort_as_half(cur.d[1])));
      const float2 d23 = __half22float2(__halves2half2(__ushort_as_half(cur.d[2]), __ushort_as_half(cur.d[3])));
#pragma unroll
      for (int xs = 0; xs < 2; ++xs) {
        int nx;
        float fx;
        axis_pt((xs ? cur.f1 : cur.f0) - cur.kx.x, P.log_scale, P.c1, P.c0, n, nx, fx);
        const float w0 = (1.f - fx) * act, w1 = fx * act;
        const bool vx0 = (unsigned)nx < (unsigned)n, vx1 = (unsigned)(nx + 1) < (unsigned)n;
        adr[xs * 2 + 0] = b_off(vx0 ? (uint32_t)((xs * TP + nx) * 4) : 80u);
        adr[xs * 2 + 1] = b_off(vx1 ? (uint32_t)((xs * TP + nx + 1) * 4) : 84u);
        val[xs * 2 + 0] = pack_f16x2(w0 * d01.x, w0 * d01.y); val[xs * 2 + 1] = pack_f16x2(w1 * d01.x, w1 * d01.y);
        hi2[xs * 2 + 0] = pack_f16x2(w0 * d23.x, w0 * d23.y); hi2[xs * 2 + 1] = pack_f16x2(w1 * d23.x, w1 * d23.y);
      }
    }

    const int round = it / STAGES;
    // (the stores depend on everything computed above, so t1 is taken after the compute phase has retired)
    const long long t1 = P.clk ? clock64() + (long long)((adr[0] & 0) + (adr[NA - 1] & 0) + (val[NA - 1] & 0)) : 0;
    if (round > 0) mbar_wait(bar_empty + stage, (round - 1) & 1);        // the MMAs that read this stage are done
    const long long t2 = P.clk ? clock64() : 0;
    if (!(P.dbg & 2)) {
      if (ROLE < 2) {
#pragma unroll
        for (int i = 0; i < NA; ++i) sts16(old[i], 0u);
#pragma unroll
        for (int i = 0; i < NA; ++i) sts16(adr[i], val[i]);
      } else {
#pragma unroll
        for (int i = 0; i < NA; ++i) sts64(old[i], 0u, 0u);
#pragma unroll
        for (int i = 0; i < NA; ++i) sts64(adr[i], val[i], hi2[i]);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_full + stage);
    const long long t3 = P.clk ? clock64() : 0;
    // reload this register set with the item two rounds ahead
    advance();
    if (it + 2 * STAGES < my_items && !(P.dbg & 4)) { load(cur, b, q, kt); ckt = kt; }
    if (P.clk) { const long long t4 = clock64(); c_comp += t1 - t0; c_wait += t2 - t1; c_store += t3 - t2; c_load += t4 - t3; }
  };
  for (int it = stage; it < my_items; it += 2 * STAGES) {
    step(r0, kt0, it, adr0, adr1);
    if (it + STAGES < my_items) step(r1, kt1, it + STAGES, adr1, adr0);
  }
  if (P.clk && threadIdx.x == 0) { atomicAdd(P.clk + 6, (unsigned long long)(tk1 - tk0)); atomicAdd(P.clk + 7, (unsigned long long)(clock64() - tk1)); }
  if (P.clk && lane == 0) {
    atomicAdd(P.clk + 0, (unsigned long long)c_comp); atomicAdd(P.clk + 1, (unsigned long long)c_wait);
    atomicAdd(P.clk + 2, (unsigned long long)c_store); atomicAdd(P.clk + 3, (unsigned long long)c_load);
  }
}

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
      for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s, WPS); mbar_init(bar_empty + s, MMA_WARPS); }
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
    const int stage = warp / WPS, pw = warp % WPS;
    if (pw < 2) produce<0>(P, smem, bar_full, bar_empty, stage, pw & 1, lane, i_begin, my_items, tk0, tk1);
    else if (pw < 4) produce<1>(P, smem, bar_full, bar_empty, stage, pw & 1, lane, i_begin, my_items, tk0, tk1);
    else produce<2>(P, smem, bar_full, bar_empty, stage, pw & 1, lane, i_begin, my_items, tk0, tk1);
  } else {
    // ------------------------------------------------------------------------------------------ MMA issuers
    // Warp PROD_WARPS + v issues the 4 K-steps of variant v of every stage.  (One lane issuing all 16 MMAs of a stage was
    // the bottleneck of the first version: ~95 cycles of descriptor moves and issue per MMA against 40 cycles of execution.)
    if (lane == 0) {
      const int v = warp - PROD_WARPS;
      const uint32_t idesc = umma_idesc_f16_major(128, NCOL, false, true);   // A K-major, B MN-major
      const uint32_t td = tmem_base + v * NCOL;
      uint64_t da[STAGES], db[STAGES];
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const uint32_t sbase = smem_u32(smem + s * STAGE_BYTES);
        da[s] = umma_desc_sw128(sbase + v * A_BYTES);
        db[s] = umma_desc_sw128_mn(sbase + 4 * A_BYTES, KS * 128, 1024);   // column blocks 8192 B apart, 8-pair groups 1024 B
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
                umma_bf16(td, da[s] + (uint64_t)(kk * 2), db[s] + (uint64_t)(kk * 128), idesc, (it0 + s) > 0 || kk > 0);   // 16 pairs: +32 B of an A row, +16 rows of B
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
