// dt6: dTables of AXIS-ALIGNED boxes as a dense tcgen05 contraction (the default dTables kernel; boxes that are not axis
// aligned go through dt3 with only_slow = 1, rpe_dtables.cu).
//
// For an axis-aligned box the trilinear weight of vertex (sx, sy, sz) in {+,-}^3 factors per axis,
//     w(q, k; z, y, x) = hz^{sz}(q,k)[z] * hy^{sy}(q,k)[y] * hx^{sx}(q,k)[x],
// every h a "hat" with (at most) two non-zero entries among the n <= 10 table points of that axis, so
//     dT_{sx,sy,sz}[z][y][x][h] = sum_pairs ( hz^{sz}[z] dS[h] ) * ( hy^{sy}[y] hx^{sx}[x] )
// is ONE GEMM over the pair index:  D [80 (sz, z, h) x 400 (sy, y, sx, x)]  +=  A [80 x pairs] * B [pairs x 400].
// The operands are 96 % zeros -- but scattering the 16 + 16 non-zero fp16 values of a pair into shared-memory tiles costs
// ~25 instructions per pair and axis pair, where sorting the pairs by table cell and accumulating them in registers (dt3)
// costs hundreds, and the tensor pipe is otherwise idle in this phase of the step.  The accumulator (400 fp32 TMEM columns
// x 128 lanes) stays in tensor memory for the whole launch: no atomics, no flushes; every CTA leaves one private copy and a
// second kernel sums the copies in a fixed order (bit-reproducible).
//
//   stage   = 64 pairs (one query x 64 consecutive keys):
//             B tile [408 rows x 64 pairs] fp16, K-major rows of 128 B with the 128-byte swizzle (the layout TMA would write);
//               row = ((sy * 10 + y) * 2 + sx) * 10 + x, rows 400..407 = dump rows for corners outside the table (the zero
//               padding of grid_sample);
//             A tile stored MN-MAJOR ([64 pairs][128 rows] in two 64-row blocks, 128-byte swizzle): row = (sz * 10 + z) * 4 + h,
//               so the 4 heads of a z point are 8 contiguous bytes of a pair's line -> ONE 64-bit store; rows 80..87 = dump.
//   warps   : 3 stages x 6 producer warps (lane = pair; a warp owns 32 pairs and one role: the B rows of y+, the B rows of
//             y-, or the A tile -- see produce<>), two MMA warps (one lane each: columns [0, 256) and [256, 400) of D,
//             4 K-steps per stage: tcgen05.mma M = 128, N = 256 / 144, K = 16).  A producer computes the next item while
//             the MMAs of its stage run, then zeroes the entries it wrote last time and writes the new ones.
//   bound   : (128 + 72) cycles per K-step x 4 = 800 cycles of tensor pipe per 64 pairs and SM; operand reads from shared
//             memory 21 KB per K-step = 96 ... 120 B/clk.
//   history : the first layout had (z, y) on the M side (four 100-row A tiles, one per (sz, sy)) and N = 80 (sx, x, h):
//             640 tensor cycles per item, but an M = 128, N = 80 MMA reads 6.5 KB of operands for 40 cycles of math -- 162 B/clk
//             against the 128 B/clk shared memory delivers -- and the producers' scattered stores compete for the same port:
//             2.0 ms per launch, MMA warps idle a third of the time.  Operand bytes per MAC fall with N, and this is the one
//             factorisation of the contraction with a large N.
#include "rpe_internal.h"
#include "rpe_fast.cuh"
#include <stdlib.h>

namespace dt6 {

using namespace tc;

constexpr int STAGES = 3, KS = 64;
constexpr int TP = 10;                                  // table points per axis the tiles are laid out for (n <= TP)
constexpr int NB = 2 * TP * 2 * TP;                     // 400 accumulator columns: (y sign, y point, x sign, x point)
constexpr int B_ROWS = NB + 8, B_BYTES = B_ROWS * 128;  // rows 400..407: dump
constexpr int MA = 2 * TP * 4;                          // 80 accumulator rows: (z sign, z point, head)
constexpr int A_BYTES = 2 * KS * 128;                   // MN-major: two 64-row blocks of [64 pairs][128 B]; rows 80..87: dump
constexpr int N0 = 256, N1 = NB - N0;                   // the two MMAs of a K-step
constexpr int STAGE_BYTES = B_BYTES + A_BYTES;          // 68608
constexpr int WPS = 6;                                  // producer warps per stage
constexpr int PROD_WARPS = STAGES * WPS;
constexpr int MMA_WARPS = 2;                            // one per column chunk of D
constexpr int THREADS = (PROD_WARPS + MMA_WARPS) * 32;
constexpr int COPY_FLOATS = MA * NB;                    // one private copy: [(sz, z, h)][(sy, y, sx, x)]
static_assert(B_BYTES % 1024 == 0 && A_BYTES % 1024 == 0 && (N0 * 128) % 1024 == 0, "tiles start on 1024-byte swizzle atoms");
static_assert(N1 % 16 == 0 && N0 % 16 == 0 && NB <= 512, "MMA shapes / TMEM columns");

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
  int n0;                                               // columns of the first MMA of a K-step (the second takes NB - n0)
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
//   ROLE 0 / 1 : the y sign sy = ROLE; writes, for x+ and x-, the 4 entries hy^{sy}[y] hx^{sx}[x] of B   (3 axes: x+, x-, y^{sy})
//   ROLE 2     : the A tile: for z+ and z-, the two z table points times the 4 heads of dS = four 64-bit stores (2 axes)
// Per step: compute the entries of the next item from registers loaded two rounds ago, wait until the MMAs that read the
// stage are done, zero the entries written last time, write the new ones, hand the stage over, reload the register set.
template <int ROLE>
__device__ __forceinline__ void produce(const Params& P, uint8_t* smem, uint64_t* bar_full, uint64_t* bar_empty, int stage, int khalf,
                                        int lane, long long i_begin, int my_items, long long tk0, long long tk1) {
  constexpr int NA = (ROLE == 2) ? 4 : 8;               // stores per step (16-bit for B, 64-bit for A)
  const uint32_t kloc = (uint32_t)(khalf * 32 + lane);
  const uint32_t lanec = ((kloc >> 3) << 4) | ((kloc & 7u) << 1);
  const uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES);
  const uint32_t tB = sbase;
  // MN-major A: this pair's 128-byte line of row block 0; the 4 heads of row group m (a multiple of 4) are 8 bytes
  const uint32_t arow = sbase + B_BYTES + kloc * 128u, asw = kloc & 7u;
  auto a_off = [&](uint32_t m) { return arow + (m >> 6) * (uint32_t)(KS * 128) + ((((m & 63u) >> 3) ^ asw) << 4) + ((m & 4u) << 1); };
  const int n = P.n;

  struct Raw {                    // what a lane loads for one pair
    float4 kx;                    // key xyz
    float f0, f1, f2;             // box faces: ROLE 0/1: x+, x-, y^{sy};  ROLE 2: z+, z-
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
    if (ROLE < 2) { w.f0 = __ldg(g); w.f1 = __ldg(g + 4); w.f2 = __ldg(g + ROLE * 4 + 1); }
    else { w.f0 = __ldg(g + 2); w.f1 = __ldg(g + 6); }
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
  for (int i = 0; i < NA; ++i) adr0[i] = adr1[i] = (ROLE == 2) ? a_off((uint32_t)MA) : tB + sw_off((uint32_t)NB, lanec);

  long long c_comp = 0, c_wait = 0, c_store = 0, c_load = 0;
  auto step = [&](Raw& cur, int& ckt, int it, uint32_t (&adr)[NA], const uint32_t (&old)[NA]) {
    const long long t0 = P.clk ? clock64() : 0;
    const bool active = (ckt * KS + (int)kloc) < P.nK && cur.fast != 0;   // inside nK, axis-aligned box
    const float act = active ? 1.f : 0.f;
    uint32_t val[NA], hi2[4];                           // hi2: heads 2, 3 of the 64-bit A stores
    if (ROLE < 2) {
      int nx[2], ny;
      float fx[2], fy;
      axis_pt(cur.f0 - cur.kx.x, P.log_scale, P.c1, P.c0, n, nx[0], fx[0]);
      axis_pt(cur.f1 - cur.kx.x, P.log_scale, P.c1, P.c0, n, nx[1], fx[1]);
      axis_pt(cur.f2 - cur.kx.y, P.log_scale, P.c1, P.c0, n, ny, fy);
      const bool vy0 = (unsigned)ny < (unsigned)n, vy1 = (unsigned)(ny + 1) < (unsigned)n;
      const float wy0 = (1.f - fy) * act, wy1 = fy * act;
#pragma unroll
      for (int xs = 0; xs < 2; ++xs) {
        const bool vx0 = (unsigned)nx[xs] < (unsigned)n, vx1 = (unsigned)(nx[xs] + 1) < (unsigned)n;
        const float wx1 = fx[xs], wx0 = 1.f - wx1;
        const int r00 = ((ROLE * TP + ny) * 2 + xs) * TP + nx[xs];          // row of (y0, x0); y + 1: + 2 * TP, x + 1: + 1
        const uint32_t p0 = pack_f16x2(wy0 * wx0, wy0 * wx1), p1 = pack_f16x2(wy1 * wx0, wy1 * wx1);
        adr[xs * 4 + 0] = tB + sw_off((vy0 && vx0) ? (uint32_t)r00 : (uint32_t)NB, lanec);
        adr[xs * 4 + 1] = tB + sw_off((vy0 && vx1) ? (uint32_t)(r00 + 1) : (uint32_t)(NB + 1), lanec);
        adr[xs * 4 + 2] = tB + sw_off((vy1 && vx0) ? (uint32_t)(r00 + 2 * TP) : (uint32_t)(NB + 2), lanec);
        adr[xs * 4 + 3] = tB + sw_off((vy1 && vx1) ? (uint32_t)(r00 + 2 * TP + 1) : (uint32_t)(NB + 3), lanec);
        val[xs * 4 + 0] = p0; val[xs * 4 + 1] = p0 >> 16; val[xs * 4 + 2] = p1; val[xs * 4 + 3] = p1 >> 16;
      }
    } else {
      // hz[z] * dS[h]: fp32 products rounded once to fp16 (dS arrives as scaled fp16)
      const float2 d01 = __half22float2(__halves2half2(__ushort_as_half(cur.d[0]), __ushort_as_half(cur.d[1])));
      const float2 d23 = __half22float2(__halves2half2(__ushort_as_half(cur.d[2]), __ushort_as_half(cur.d[3])));
#pragma unroll
      for (int zs = 0; zs < 2; ++zs) {
        int nz;
        float fz;
        axis_pt((zs ? cur.f1 : cur.f0) - cur.kx.z, P.log_scale, P.c1, P.c0, n, nz, fz);
        const float w0 = (1.f - fz) * act, w1 = fz * act;
        const bool vz0 = (unsigned)nz < (unsigned)n, vz1 = (unsigned)(nz + 1) < (unsigned)n;
        adr[zs * 2 + 0] = a_off(vz0 ? (uint32_t)((zs * TP + nz) * 4) : (uint32_t)MA);
        adr[zs * 2 + 1] = a_off(vz1 ? (uint32_t)((zs * TP + nz + 1) * 4) : (uint32_t)(MA + 4));
        val[zs * 2 + 0] = pack_f16x2(w0 * d01.x, w0 * d01.y); val[zs * 2 + 1] = pack_f16x2(w1 * d01.x, w1 * d01.y);
        hi2[zs * 2 + 0] = pack_f16x2(w0 * d23.x, w0 * d23.y); hi2[zs * 2 + 1] = pack_f16x2(w1 * d23.x, w1 * d23.y);
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
    // Warp PROD_WARPS + c issues the 4 K-steps of column chunk c of every stage.
    if (lane == 0) {
      const int c = warp - PROD_WARPS;
      const uint32_t idesc = umma_idesc_f16_major(128, c ? NB - P.n0 : P.n0, true, false);   // A MN-major, B K-major
      const uint32_t td = tmem_base + c * P.n0;
      uint64_t da[STAGES], db[STAGES];
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const uint32_t sbase = smem_u32(smem + s * STAGE_BYTES);
        da[s] = umma_desc_sw128_mn(sbase + B_BYTES, KS * 128, 1024);    // row blocks 8192 B apart, 8-pair groups 1024 B
        db[s] = umma_desc_sw128(sbase + c * P.n0 * 128);
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
              for (int kk = 0; kk < KS / 16; ++kk)     // 16 pairs: +16 lines of A, +32 B of a B row
                umma_bf16(td, da[s] + (uint64_t)(kk * 128), db[s] + (uint64_t)(kk * 2), idesc, (it0 + s) > 0 || kk > 0);
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
  if (warp < 12) {                             // 3 column groups x 4 lane quarters; rows >= 80 of D are never read
    mbar_wait_relaxed(bar_done, 0);
    tc_fence_after();
    const int quad = warp & 3, row = quad * 32 + lane;
    float* dst = P.priv + (size_t)blockIdx.x * COPY_FLOATS;
#pragma unroll 1
    for (int c = warp >> 2; c < NB / 16; c += 3) {
      uint32_t rr[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 16), rr);
      tmem_ld_wait();
      if (row < MA) {
        float4* o = reinterpret_cast<float4*>(dst + (size_t)row * NB + c * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PROD_WARPS) tmem_dealloc<512>(tmem_base);
}

// out[vertex][z][y][x][h] (+)= (1 / scale) * sum over the private copies, in a fixed order.  A CTA owns 32 consecutive
// elements of a copy (coalesced reads); its 8 warps split the copies; every element maps to one output.
__global__ void __launch_bounds__(256) rpe_dtables_umma_reduce_kernel(const float* __restrict__ priv, int copies, int n,
                                                                      float* __restrict__ out, const unsigned* absmax_bits, int dense,
                                                                      int accumulate) {
  __shared__ float part[8][32];
  const float inv = 1.0f / vdetr_dt_scale(*absmax_bits, dense);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = blockIdx.x * 32; i0 < COPY_FLOATS; i0 += gridDim.x * 32) {
    const int src = i0 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (src < COPY_FLOATS) {
      int c = warp;
      for (; c + 24 < copies; c += 32) {
        s0 += priv[(size_t)c * COPY_FLOATS + src]; s1 += priv[(size_t)(c + 8) * COPY_FLOATS + src];
        s2 += priv[(size_t)(c + 16) * COPY_FLOATS + src]; s3 += priv[(size_t)(c + 24) * COPY_FLOATS + src];
      }
      for (; c < copies; c += 8) s0 += priv[(size_t)c * COPY_FLOATS + src];
    }
    part[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && src < COPY_FLOATS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][lane];
      // src = m * 400 + col, m = (sz * 10 + z) * 4 + h, col = ((sy * 10 + y) * 2 + sx) * 10 + x
      const int m = src / NB, col = src - m * NB;
      const int h = m & 3, z = (m >> 2) % TP, zs = (m >> 2) / TP;
      const int x = col % TP, xs = (col / TP) & 1, y = (col / (2 * TP)) % TP, ys = col / (2 * TP * TP);
      if (z < n && y < n && x < n) {
        // vertex sign table: 0:(+,+,-) 1:(+,-,-) 2:(-,-,-) 3:(-,+,-) 4..7: the same with z +
        const int v = (zs ? 0 : 4) + (xs ? (ys ? 2 : 3) : (ys ? 1 : 0));
        const size_t o = ((((size_t)v * n + z) * n + y) * n + x) * 4 + h;
        out[o] = (accumulate ? out[o] : 0.f) + t * inv;
      }
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
  P.n0 = dt6::N0;
  { const char* e = getenv("VDETR_DT6_N0"); if (e && atoi(e) >= 144 && atoi(e) <= 256 && atoi(e) % 16 == 0) P.n0 = atoi(e); }   // developer
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
  dt6::rpe_dtables_umma_reduce_kernel<<<(dt6::COPY_FLOATS + 31) / 32, 256, 0, st>>>(priv, grid, n, dtables, absmax_bits, dense_scale, accumulate);
  VDETR_LAUNCH_CHECK();
  return 0;
}
