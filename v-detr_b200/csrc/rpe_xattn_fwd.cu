// Fused Vertex-RPE attention forward for sm_100a (impl = 0, the product path).
//
//   O = softmax_k( Q K^T + rpe(ref_pts, xyz, tables) ) V          /root/reference/models/vdetr_transformer.py:708-753
//
// One persistent CTA per SM.  A work item is (scene, 128-row query tile, key split):
//   * MQA (cross attention, ShareSelfAttention): the 4 query heads share K/V, so the 128 MMA rows are
//     32 queries x 4 heads and ONE geometry evaluation per (query, key) pair feeds all four heads (float4 cell);
//   * MHA (decoder self attention): 128 queries of one head, no bias.
// Per 64-key tile:
//   TMA      K tile [64 x 64] fp16 hi + lo (128B swizzle), V^T tile [64 d x 64 keys], key xyz (1-D bulk copy)
//   tcgen05  S = Qh Kh^T + Ql Kh^T + Qh Kl^T  -> TMEM (double buffered), issued one tile ahead by the control warp;
//            q = Qh + Ql, k = Kh + Kl are fp16 hi/lo splits of the fp32 operands, so S carries ~2^-22 relative
//            rounding instead of fp16's 2^-11 (the tensor pipe is idle: three MMAs cost nothing here)
//   CUDA     all 16 compute warps: Vertex-RPE bias of the 32 x 64 (query,key) pairs (lanes = 32 consecutive
//            keys, so table reads broadcast), staged through shared memory;
//            then every thread owns (row, 16 key columns): tcgen05.ld S, + bias, online softmax (row max via a
//            4-way smem exchange), P -> shared memory in the UMMA K-major swizzled layout
//   tcgen05  O += P V   (accumulator stays in TMEM for the whole item; rescaled in place when the max moves)
//   dropout  (training, nn.Dropout on the probabilities, vdetr_transformer.py:751-752) Philox keep-mask per (row, key)
//            applied to P before the PV product; the softmax denominator uses the undropped P
// Nothing of size nQ x nK ever reaches global memory.
#include "rpe_internal.h"
#include "tc_common.cuh"
#include "rpe_fast.cuh"
#include "philox.cuh"

namespace {

using namespace tc;
using rpe::rpe_bias_pair;

constexpr int NCOMPUTE_WARPS = 16;
constexpr int NCOMPUTE = NCOMPUTE_WARPS * 32;     // 512
constexpr int NTHREADS = NCOMPUTE + 32;           // + control warp
constexpr int BM = 128;                           // MMA rows per item
constexpr int BN = 64;                            // keys per tile
constexpr int HD = 64;
constexpr int QT = 32;                            // queries per MQA tile
constexpr int GEO_F4 = 9;                         // float4 per query geometry record
constexpr int BIAS_STRIDE_F4 = BN + 1;            // padded row (per query) of the bias staging buffer
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int TMEM_COLS = 256;                    // S0 [0,64) S1 [64,128) O [128,192)

struct FwdParams {
  int B, nQ, nK, nQp, nKp, kvh;
  int mtiles;             // per scene
  int splits, tiles_per_split;
  int items;              // B * mtiles * splits
  int grid_n;
  float log_scale, c1, c0;      // p = copysign(lg2(|d|*ls+1) * c1, d) + c0
  const float4* xyz4;           // [B][nKp]
  const float4* geo;            // [B][nQp][GEO_F4]
  const float4* tables;         // [8][n^3]
  float* out;                   // [B,nQ,4,64]
  float* lse;                   // [B,4,nQ]
  float* part_o;                // [items][128][64]   (splits > 1)
  float2* part_ml;              // [items][128] (m in log2 units, l)
  float4* bias_out;             // [B][nQp][nKp] bias of the 4 heads per pair, or null (saved for the backward)
  const unsigned long long* drop_seed;   // device pointer to the 64-bit dropout seed, or null (no dropout)
  uint32_t drop_thresh;         // philox::thresh_of(p)
  float drop_inv_keep;          // 1 / (1 - p)
};

struct SmemLayout {
  uint32_t tables, q, ql, k, kl, vt, p, bias, xyz, geo, smax, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int table_bytes) {
  SmemLayout L;
  uint32_t o = 0;
  L.q = o;      o += BM * 128;                       // 16 KB  (1024-aligned: o == 0)
  L.ql = o;     o += BM * 128;                       // 16 KB  rounding residual of Q
  L.k = o;      o += BN * 128;                       //  8 KB
  L.kl = o;     o += BN * 128;                       //  8 KB  rounding residual of K
  L.vt = o;     o += HD * 128;                       //  8 KB
  L.p = o;      o += BM * 128;                       // 16 KB
  L.tables = o; o += (uint32_t)((table_bytes + 15) / 16 * 16);
  L.bias = o;   o += QT * BIAS_STRIDE_F4 * 16;       // 33,280 B
  L.xyz = o;    o += 2 * BN * 16;
  L.geo = o;    o += QT * GEO_F4 * 16;
  L.smax = o;   o += BM * 4 * 4;
  L.bars = o;   o += 128;
  L.total = o;
  return L;
}

// ------------------------------------------------------------------------------------------------ the kernel
template <bool HAS_BIAS, bool MQA>
__global__ void __launch_bounds__(NTHREADS, 1)
rpe_xattn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQl,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmKl,
                     const __grid_constant__ CUtensorMap tmVt, const FwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B tiles need 1024-B alignment
  const SmemLayout L = smem_layout(HAS_BIAS ? rpe::pair_table_bytes(P.grid_n) : 0);
  uint8_t* sQ = smem + L.q;
  uint8_t* sQl = smem + L.ql;
  uint8_t* sK = smem + L.k;
  uint8_t* sKl = smem + L.kl;
  uint8_t* sVt = smem + L.vt;
  uint8_t* sP = smem + L.p;
  const char* sTab = reinterpret_cast<const char*>(smem + L.tables);
  float4* sBias = reinterpret_cast<float4*>(smem + L.bias);
  float4* sXyz = reinterpret_cast<float4*>(smem + L.xyz);
  float4* sGeo = reinterpret_cast<float4*>(smem + L.geo);
  float* sMax = reinterpret_cast<float*>(smem + L.smax);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;       // [2] one per key-xyz buffer: waiters can never fall two phases behind
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_kfree = bars + 4;
  uint64_t* bar_s = bars + 5;       // [2]
  uint64_t* bar_p = bars + 7;
  uint64_t* bar_pv = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_control = warp == NCOMPUTE_WARPS;

  if (is_control) {
    if (lane == 0) {
      mbar_init(bar_q, 1); mbar_init(bar_k + 0, 1); mbar_init(bar_k + 1, 1); mbar_init(bar_v, 1); mbar_init(bar_kfree, 1);
      mbar_init(bar_s + 0, 1); mbar_init(bar_s + 1, 1);
      mbar_init(bar_p, NCOMPUTE); mbar_init(bar_pv, 1);
      fence_barrier_init();
      prefetch_tmap(&tmQ); prefetch_tmap(&tmQl); prefetch_tmap(&tmK); prefetch_tmap(&tmKl); prefetch_tmap(&tmVt);
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(tmem_slot);
  } else if (HAS_BIAS) {
    // tables -> shared memory as fp16 x-pairs, once per (persistent) CTA
    rpe::load_pair_tables(reinterpret_cast<uint4*>(smem + L.tables), P.tables, P.grid_n, tid, NCOMPUTE);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 128;

  const uint32_t idesc_s = umma_idesc_f16(BM, BN);
  const uint32_t idesc_o = umma_idesc_f16(BM, HD);
  philox::Dropout drop;
  drop.thresh = P.drop_seed ? P.drop_thresh : 0u;
  drop.inv_keep = P.drop_inv_keep;
  {
    const unsigned long long seed = P.drop_seed ? __ldg(P.drop_seed) : 0ull;
    drop.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  }
  // S(tile) = Qh Kh^T + Ql Kh^T + Qh Kl^T into TMEM columns [tcol, tcol + BN)
  auto issue_s = [&](uint32_t tcol) {
    const uint64_t dqh = umma_desc_sw128(smem_u32(sQ)), dql = umma_desc_sw128(smem_u32(sQl));
    const uint64_t dkh = umma_desc_sw128(smem_u32(sK)), dkl = umma_desc_sw128(smem_u32(sKl));
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) umma_bf16(tcol, dqh + (uint64_t)(kk * 2), dkh + (uint64_t)(kk * 2), idesc_s, kk > 0);
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) umma_bf16(tcol, dql + (uint64_t)(kk * 2), dkh + (uint64_t)(kk * 2), idesc_s, true);
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) umma_bf16(tcol, dqh + (uint64_t)(kk * 2), dkl + (uint64_t)(kk * 2), idesc_s, true);
  };

  uint32_t g = 0;          // tiles processed by this CTA so far (phase bookkeeping)
  uint32_t it = 0;         // items processed by this CTA so far

  for (int item = blockIdx.x; item < P.items; item += gridDim.x, ++it) {
    const int split = item % P.splits;
    const int mt = (item / P.splits) % P.mtiles;
    const int b = item / (P.splits * P.mtiles);
    const int tile_begin = split * P.tiles_per_split;
    const int total_tiles = P.nKp / BN;
    const int T = min(P.tiles_per_split, total_tiles - tile_begin);
    int q0, hsel;
    if (MQA) { q0 = mt * QT; hsel = 0; } else { q0 = (mt >> 2) * BM; hsel = mt & 3; }
    const int hk = MQA ? 0 : hsel;
    const int qrow0 = MQA ? (b * P.nQp + q0) * 4 : ((b * 4 + hsel) * P.nQp + q0);
    const int krow0 = (b * P.kvh + hk) * P.nKp;
    const int vrow0 = (b * P.kvh + hk) * HD;

    if (is_control) {
      // ======================================================================== control warp (one lane)
      if (lane == 0) {
        if (it > 0) mbar_wait(bar_pv, (g - 1) & 1);       // last PV of the previous item: sVt / sQ are free
        mbar_arrive_expect_tx(bar_q, 2 * BM * 128);
        tma_load_2d(sQ, &tmQ, 0, qrow0, bar_q);
        tma_load_2d(sQl, &tmQl, 0, qrow0, bar_q);
        mbar_arrive_expect_tx(bar_k + (g & 1), 2 * BN * 128 + (HAS_BIAS ? BN * 16 : 0));
        tma_load_2d(sK, &tmK, 0, krow0 + tile_begin * BN, bar_k + (g & 1));
        tma_load_2d(sKl, &tmKl, 0, krow0 + tile_begin * BN, bar_k + (g & 1));
        if (HAS_BIAS)
          bulk_load_1d(sXyz + (g & 1) * BN, P.xyz4 + (size_t)b * P.nKp + tile_begin * BN, BN * 16, bar_k + (g & 1));
        mbar_arrive_expect_tx(bar_v, HD * 128);
        tma_load_2d(sVt, &tmVt, tile_begin * BN, vrow0, bar_v);
        mbar_wait(bar_q, it & 1);
        mbar_wait(bar_k + (g & 1), (g >> 1) & 1);
        tc_fence_after();
        issue_s(tS0 + (g & 1) * BN);
        umma_commit(bar_s + (g & 1));
        umma_commit(bar_kfree);
        for (int j = 0; j < T; ++j) {
          const uint32_t gj = g + j;
          if (j + 1 < T) {
            mbar_wait(bar_kfree, gj & 1);
            uint64_t* bk = bar_k + ((gj + 1) & 1);
            mbar_arrive_expect_tx(bk, 2 * BN * 128 + (HAS_BIAS ? BN * 16 : 0));
            tma_load_2d(sK, &tmK, 0, krow0 + (tile_begin + j + 1) * BN, bk);
            tma_load_2d(sKl, &tmKl, 0, krow0 + (tile_begin + j + 1) * BN, bk);
            if (HAS_BIAS)
              bulk_load_1d(sXyz + ((gj + 1) & 1) * BN, P.xyz4 + (size_t)b * P.nKp + (tile_begin + j + 1) * BN, BN * 16, bk);
            mbar_wait(bk, ((gj + 1) >> 1) & 1);
            tc_fence_after();
            issue_s(tS0 + ((gj + 1) & 1) * BN);
            umma_commit(bar_s + ((gj + 1) & 1));
            umma_commit(bar_kfree);
          }
          mbar_wait_relaxed(bar_p, gj & 1);                  // a whole tile of bias + softmax work away
          mbar_wait(bar_v, gj & 1);
          tc_fence_after();
          {
            const uint64_t da = umma_desc_sw128(smem_u32(sP)), db = umma_desc_sw128(smem_u32(sVt));
#pragma unroll
            for (int kk = 0; kk < BN / 16; ++kk)
              umma_bf16(tO, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc_o, (j > 0) || (kk > 0));
            umma_commit(bar_pv);
          }
          if (j + 1 < T) {
            mbar_wait(bar_pv, gj & 1);
            mbar_arrive_expect_tx(bar_v, HD * 128);
            tma_load_2d(sVt, &tmVt, (tile_begin + j + 1) * BN, vrow0, bar_v);
          }
        }
      }
      __syncwarp();
    } else {
      // ======================================================================== compute warps
      const int quarter = warp & 3, slice = warp >> 2;
      const int row = quarter * 32 + lane;                 // TMEM lane == MMA row
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      if (HAS_BIAS) {
        // query geometry records of this tile -> smem (previous item's readers are past their last barrier)
        const float4* src = P.geo + ((size_t)b * P.nQp + q0) * GEO_F4;
        for (int i = tid; i < QT * GEO_F4; i += NCOMPUTE) sGeo[i] = __ldg(src + i);
        named_bar_sync(1, NCOMPUTE);
      }
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < T; ++j) {
        const uint32_t gj = g + j;
        const int key0 = (tile_begin + j) * BN;
        if (HAS_BIAS) {
          mbar_wait(bar_k + (gj & 1), (gj >> 1) & 1);        // key xyz of this tile has landed
          const int kg = warp & 1;                           // key group (32 consecutive keys), lane = key
          const float4 kx = sXyz[(gj & 1) * BN + kg * 32 + lane];
#pragma unroll 1
          for (int u = 0; u < 4; ++u) {
            const int q = (warp >> 1) + 8 * u;
            const float4 bias = rpe_bias_pair(sGeo + q * GEO_F4, kx.x, kx.y, kx.z, sTab, P.grid_n, P.log_scale, P.c1, P.c0);
            sBias[q * BIAS_STRIDE_F4 + kg * 32 + lane] = bias;
            if (P.bias_out) P.bias_out[((size_t)b * P.nQp + q0 + q) * P.nKp + key0 + kg * 32 + lane] = bias;
          }
        }
        named_bar_sync(1, NCOMPUTE);                         // (a) bias tile complete (and previous smax reads done)
        mbar_wait(bar_s + (gj & 1), (gj >> 1) & 1);          // S tile is in TMEM
        tc_fence_after();
        uint32_t sr[16];
        tmem_ld16(tS0 + (gj & 1) * BN + lane_addr + slice * 16, sr);
        tmem_ld_wait();
        float x[16];
        float pmax = -INFINITY;
        {
          const bool tail_tile = key0 + BN > P.nK;
          const float* brow = nullptr;
          if (HAS_BIAS) brow = reinterpret_cast<const float*>(sBias + (row >> 2) * BIAS_STRIDE_F4 + slice * 16) + (row & 3);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float s = __uint_as_float(sr[c]);
            if (HAS_BIAS) s += brow[c * 4];
            s *= LOG2E;
            if (tail_tile && key0 + slice * 16 + c >= P.nK) s = -INFINITY;       // only the last key tile is ragged
            x[c] = s;
            pmax = fmaxf(pmax, s);
          }
        }
        sMax[row * 4 + slice] = pmax;
        named_bar_sync(2, NCOMPUTE);                         // (b)
        const float4 pm = *reinterpret_cast<const float4*>(sMax + row * 4);
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(pm.x, pm.y), fmaxf(pm.z, pm.w)));
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = ex2_approx(m_run - m_use);       // m_run = -inf -> 0
        float psum = 0.f;
        uint32_t pk[8];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const float p0 = ex2_approx(x[c] - m_use), p1 = ex2_approx(x[c + 1] - m_use);
          pk[c >> 1] = pack_f16x2(p0, p1);
          // accumulate the row sum from the fp16-rounded values that the PV MMA will actually use
          const float2 rb = __half22float2(*reinterpret_cast<const __half2*>(&pk[c >> 1]));
          psum += rb.x + rb.y;
        }
        if (drop.thresh) {
          // dropout on the probabilities: the PV operand becomes P * keep / (1 - p); the denominator above does not change
          const uint32_t keep = philox::keep_mask16(drop, (uint32_t)(qrow0 + row), (uint32_t)((key0 >> 4) + slice));
          const __half hk = __float2half_rn(drop.inv_keep), hz = __float2half_rn(0.f);
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const __half2 sc = __halves2half2(((keep >> c) & 1u) ? hk : hz, ((keep >> (c + 1)) & 1u) ? hk : hz);
            const __half2 v = __hmul2(*reinterpret_cast<const __half2*>(&pk[c >> 1]), sc);
            pk[c >> 1] = *reinterpret_cast<const uint32_t*>(&v);
          }
        }
        l_run = l_run * alpha + psum;
        m_run = m_new;
        if (j > 0) {
          mbar_wait(bar_pv, (gj - 1) & 1);                   // PV of the previous tile done: O readable, sP free
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {      // rescale this thread's 16 O columns in place
            uint32_t o[16];
            tmem_ld16(tO + lane_addr + slice * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st16(tO + lane_addr + slice * 16, o);
            tmem_st_wait();
          }
        }
        {
          // P[row][16*slice .. +15] -> sP, K-major rows of 128 B with the 128-byte swizzle (chunk ^= row % 8)
          uint8_t* prow = sP + (row >> 3) * 1024 + (row & 7) * 128;
          const int ch0 = slice * 2, ch1 = slice * 2 + 1;
          *reinterpret_cast<uint4*>(prow + ((ch0 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(prow + ((ch1 ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_p);
      }
      // ------------------------------------------------------------------ item epilogue
      mbar_wait(bar_pv, (g + T - 1) & 1);
      tc_fence_after();
      sMax[row * 4 + slice] = l_run;
      named_bar_sync(2, NCOMPUTE);
      const float4 pl = *reinterpret_cast<const float4*>(sMax + row * 4);
      const float l_tot = (pl.x + pl.y) + (pl.z + pl.w);
      uint32_t o[16];
      tmem_ld16(tO + lane_addr + slice * 16, o);
      tmem_ld_wait();
      const int q = MQA ? q0 + (row >> 2) : q0 + row;
      const int h = MQA ? (row & 3) : hsel;
      if (P.splits == 1) {
        if (q < P.nQ) {
          const float inv = 1.0f / l_tot;
          float4* dst = reinterpret_cast<float4*>(P.out + (((size_t)b * P.nQ + q) * 4 + h) * HD + slice * 16);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            dst[c] = make_float4(__uint_as_float(o[4 * c]) * inv, __uint_as_float(o[4 * c + 1]) * inv,
                                 __uint_as_float(o[4 * c + 2]) * inv, __uint_as_float(o[4 * c + 3]) * inv);
          if (slice == 0) P.lse[((size_t)b * 4 + h) * P.nQ + q] = (m_run + log2f(l_tot)) * LN2;
        }
      } else {
        float4* dst = reinterpret_cast<float4*>(P.part_o + ((size_t)item * BM + row) * HD + slice * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_float4(__uint_as_float(o[4 * c]), __uint_as_float(o[4 * c + 1]), __uint_as_float(o[4 * c + 2]),
                               __uint_as_float(o[4 * c + 3]));
        if (slice == 0) P.part_ml[(size_t)item * BM + row] = make_float2(m_run, l_tot);
      }
      tc_fence_before();
      named_bar_sync(1, NCOMPUTE);     // sMax / sGeo / TMEM O reads are done before the next item starts
    }
    g += (uint32_t)T;
  }

  tc_fence_before();
  __syncthreads();
  if (is_control) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ split combine
__global__ void fwd_combine_kernel(FwdParams P, int mqa) {
  // one thread per (b, mtile, row, 16-column slice)
  const size_t total = (size_t)P.B * P.mtiles * BM * 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int slice = (int)(i & 3);
    const int row = (int)((i >> 2) % BM);
    const size_t bm = (i >> 2) / BM;                        // b * mtiles + mt
    const int mt = (int)(bm % P.mtiles), b = (int)(bm / P.mtiles);
    int q, h;
    if (mqa) { q = mt * QT + (row >> 2); h = row & 3; } else { q = (mt >> 2) * BM + row; h = mt & 3; }
    if (q >= P.nQ) continue;
    float m = -INFINITY;
    for (int s = 0; s < P.splits; ++s) m = fmaxf(m, P.part_ml[(bm * P.splits + s) * BM + row].x);
    float lsum = 0.f, acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    for (int s = 0; s < P.splits; ++s) {
      const float2 ml = P.part_ml[(bm * P.splits + s) * BM + row];
      const float w = exp2f(ml.x - m);
      lsum += ml.y * w;
      const float4* src = reinterpret_cast<const float4*>(P.part_o + ((bm * P.splits + s) * BM + row) * HD + slice * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v = src[c];
        acc[4 * c] += v.x * w; acc[4 * c + 1] += v.y * w; acc[4 * c + 2] += v.z * w; acc[4 * c + 3] += v.w * w;
      }
    }
    const float inv = 1.0f / lsum;
    float4* dst = reinterpret_cast<float4*>(P.out + (((size_t)b * P.nQ + q) * 4 + h) * HD + slice * 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv, acc[4 * c + 3] * inv);
    if (slice == 0) P.lse[((size_t)b * 4 + h) * P.nQ + q] = (m + log2f(lsum)) * LN2;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ packing (shared with bwd)
__global__ void vdetr_absmax_kernel(const float* x, size_t n, unsigned* out_bits) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));      // non-negative floats order like uints
}

// fp16 hi/lo split of an fp32 value: x ~= hi + lo to ~2^-22 relative.  Saturated at the fp16 range (the reference is
// fp32 and would stay finite; a saturated operand is wrong but finite, an infinite one poisons the whole softmax row).
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  const float xs = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}

__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  __half h0, h1, h2, h3, l0, l1, l2, l3;
  split_f16(v.x, h0, l0); split_f16(v.y, h1, l1); split_f16(v.z, h2, l2); split_f16(v.w, h3, l3);
  const __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3), la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&la), *reinterpret_cast<const uint32_t*>(&lb));
}
__device__ __forceinline__ uint2 sat4(const float4 v, float s) {
  const __half2 a = __floats2half2_rn(fminf(fmaxf(v.x * s, -65504.f), 65504.f), fminf(fmaxf(v.y * s, -65504.f), 65504.f));
  const __half2 b = __floats2half2_rn(fminf(fmaxf(v.z * s, -65504.f), 65504.f), fminf(fmaxf(v.w * s, -65504.f), 65504.f));
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// Row operands (Q / dO, K / row-major V), key xyz and query geometry.  One thread per group of 4 consecutive channels:
// float4 loads, 8-byte fp16 stores.  (V^T has its own kernel: it is a transposition.)
__global__ void vdetr_pack_kernel(VdetrPack K) {
  const float gscale = K.dout ? vdetr_grad_scale(*K.dout_absmax_bits) : 1.f;
  const size_t nq = K.qp ? (size_t)K.B * K.nQp * 4 * 16 : 0;           // float4 groups of Qp
  const size_t nk = K.kp ? (size_t)K.B * K.kvh * K.nKp * 16 : 0;
  const size_t nx = K.has_bias ? (size_t)K.B * K.nKp : 0;
  const size_t ng = K.has_bias ? (size_t)K.B * K.nQp : 0;
  const size_t total = nq + nk + nx + ng;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (i < nq) {
      // destination row-major [rows][64]; MQA rows = (b*nQp + q)*4 + h ; MHA rows = (b*4 + h)*nQp + q
      const int d4 = (int)(i & 15);
      const size_t r = i >> 4;
      int b, q, h;
      if (K.kvh == 1) { h = (int)(r & 3); q = (int)((r >> 2) % K.nQp); b = (int)((r >> 2) / K.nQp); }
      else { q = (int)(r % K.nQp); h = (int)((r / K.nQp) & 3); b = (int)(r / ((size_t)K.nQp * 4)); }
      const size_t src = ((((size_t)b * K.nQ + q) * 4 + h) * 64) / 4 + d4;
      const float4 val = q < K.nQ ? __ldg(reinterpret_cast<const float4*>(K.q) + src) : zero4;
      uint2 hi, lo;
      split4(val, hi, lo);
      reinterpret_cast<uint2*>(K.qp)[i] = hi;
      if (K.qpl) reinterpret_cast<uint2*>(K.qpl)[i] = lo;
      if (K.dout) {
        const float4 dv = q < K.nQ ? __ldg(reinterpret_cast<const float4*>(K.dout) + src) : zero4;
        reinterpret_cast<uint2*>(K.dop)[i] = sat4(dv, gscale);
      }
    } else if (i < nq + nk) {
      const size_t e = i - nq;                               // Kp [b][hk][key][64]
      const int d4 = (int)(e & 15);
      const size_t r = e >> 4;
      const int key = (int)(r % K.nKp);
      const int hk = (int)((r / K.nKp) % K.kvh), b = (int)(r / ((size_t)K.nKp * K.kvh));
      const size_t src = ((((size_t)b * K.nK + key) * K.kvh + hk) * 64) / 4 + d4;
      const float4 val = key < K.nK ? __ldg(reinterpret_cast<const float4*>(K.k) + src) : zero4;
      uint2 hi, lo;
      split4(val, hi, lo);
      reinterpret_cast<uint2*>(K.kp)[e] = hi;
      if (K.kpl) reinterpret_cast<uint2*>(K.kpl)[e] = lo;
      if (K.vp) {                                            // row-major V as well (backward: dP = dO V^T)
        const float4 vv = key < K.nK ? __ldg(reinterpret_cast<const float4*>(K.v) + src) : zero4;
        reinterpret_cast<uint2*>(K.vp)[e] = sat4(vv, 1.f);
      }
    } else if (i < nq + nk + nx) {
      const size_t e = i - nq - nk;
      const int key = (int)(e % K.nKp), b = (int)(e / K.nKp);
      float4 o = make_float4(1e9f, 1e9f, 1e9f, 0.f);
      if (key < K.nK) {
        const float* s = K.xyz + ((size_t)b * K.nK + key) * 3;
        o = make_float4(s[0], s[1], s[2], 0.f);
      }
      K.xyz4[e] = o;
    } else {
      const size_t e = i - nq - nk - nx;
      const int q = (int)(e % K.nQp), b = (int)(e / K.nQp);
      float v[24];
#pragma unroll
      for (int t = 0; t < 24; ++t) v[t] = 0.f;
      float c = 1.f, s = 0.f;
      if (q < K.nQ) {
        const float* src = K.ref + ((size_t)b * K.nQ + q) * 24;
#pragma unroll
        for (int t = 0; t < 24; ++t) v[t] = src[t];
        if (K.ang) { const float a = K.ang[(size_t)b * K.nQ + q]; c = cosf(a); s = sinf(a); }
      }
      // axis aligned?  vertex i = centre + (sx,sy,sz) * half (SURVEY Appendix A sign table)
      const float xp = v[0], yp = v[1], zm = v[2], ym = v[4], xm = v[6], zp = v[14];
      bool fast = (K.ang == nullptr);
      const float sxp[8] = {1, 1, 0, 0, 1, 1, 0, 0}, syp[8] = {1, 0, 0, 1, 1, 0, 0, 1}, szp[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        fast = fast && (v[t * 3 + 0] == (sxp[t] != 0.f ? xp : xm)) && (v[t * 3 + 1] == (syp[t] != 0.f ? yp : ym)) &&
               (v[t * 3 + 2] == (szp[t] != 0.f ? zp : zm));
      }
      float4* dst = K.geo + e * 9;
      dst[0] = make_float4(xp, yp, zp, __int_as_float(fast ? 1 : 0));
      dst[1] = make_float4(xm, ym, zm, 0.f);
#pragma unroll
      for (int t = 0; t < 6; ++t) dst[2 + t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
      dst[8] = make_float4(c, s, 0.f, 0.f);
    }
  }
}

// V^T operand of the forward's P V product: v [B][nK][kvh][64] f32 -> vtp [B][kvh][64][nKp] fp16 (keys contiguous, zero padded).
// A CTA transposes a 64-key x 64-channel tile through shared memory: coalesced float4 reads along the channels, coalesced
// half2 stores along the keys (a thread-per-output-element kernel reads V with a 256-byte stride).
__global__ void __launch_bounds__(256) vdetr_pack_vt_kernel(VdetrPack K) {
  __shared__ float tile[64][65];
  const int ktiles = K.nKp / 64;
  const int t = blockIdx.x % ktiles, bh = blockIdx.x / ktiles;
  const int hk = bh % K.kvh, b = bh / K.kvh;
  const int key0 = t * 64;
  for (int i = threadIdx.x; i < 64 * 16; i += 256) {
    const int kl = i >> 4, d4 = i & 15, key = key0 + kl;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (key < K.nK) v = __ldg(reinterpret_cast<const float4*>(K.v + (((size_t)b * K.nK + key) * K.kvh + hk) * 64) + d4);
    tile[kl][d4 * 4 + 0] = v.x; tile[kl][d4 * 4 + 1] = v.y; tile[kl][d4 * 4 + 2] = v.z; tile[kl][d4 * 4 + 3] = v.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    const int d = i >> 5, k2 = (i & 31) * 2;
    const float a = fminf(fmaxf(tile[k2][d], -65504.f), 65504.f), c = fminf(fmaxf(tile[k2 + 1][d], -65504.f), 65504.f);
    *reinterpret_cast<__half2*>(K.vtp + (((size_t)b * K.kvh + hk) * 64 + d) * K.nKp + key0 + k2) = __floats2half2_rn(a, c);
  }
}

// ------------------------------------------------------------------------------------------------ host side
namespace {

struct FwdPlan {
  int nQp, nKp, mtiles, splits, tiles_per_split, items;
  size_t off_qp, off_qpl, off_kp, off_kpl, off_vtp, off_xyz, off_geo, off_po, off_pml, total;
};

FwdPlan make_plan(const VdetrXattnShape* s) {
  FwdPlan p;
  const bool mqa = s->kv_heads == 1;
  p.nQp = mqa ? (s->nQ + QT - 1) / QT * QT : (s->nQ + BM - 1) / BM * BM;
  p.nKp = (s->nK + BN - 1) / BN * BN;
  p.mtiles = mqa ? p.nQp / QT : (p.nQp / BM) * 4;
  const int ktiles = p.nKp / BN;
  const int sms = vdetr_num_sms();
  const long base = (long)s->B * p.mtiles;
  int best = 1;
  double best_eff = -1.0;
  for (int sp = 1; sp <= 16; ++sp) {
    if (sp > 1 && ktiles / sp < 4) break;                 // keep >= 4 key tiles per split
    const int tps = (ktiles + sp - 1) / sp;
    const int real = (ktiles + tps - 1) / tps;
    if (real != sp) continue;
    const double waves = (double)(base * sp) / sms;
    const double eff = waves / (double)((long)((base * sp + sms - 1) / sms));
    if (eff > best_eff + 0.04) { best_eff = eff; best = sp; }
    if (eff >= 0.93) { best = sp; break; }
  }
  p.splits = best;
  p.tiles_per_split = (ktiles + best - 1) / best;
  p.items = (int)(base * best);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += vdetr_align_up(bytes, 1024); return r; };
  p.off_qp = take((size_t)s->B * p.nQp * 4 * 64 * 2);
  p.off_qpl = take((size_t)s->B * p.nQp * 4 * 64 * 2);
  p.off_kp = take((size_t)s->B * s->kv_heads * p.nKp * 64 * 2);
  p.off_kpl = take((size_t)s->B * s->kv_heads * p.nKp * 64 * 2);
  p.off_vtp = take((size_t)s->B * s->kv_heads * p.nKp * 64 * 2);
  p.off_xyz = take(s->has_bias ? (size_t)s->B * p.nKp * 16 : 0);
  p.off_geo = take(s->has_bias ? (size_t)s->B * p.nQp * GEO_F4 * 16 : 0);
  p.off_po = take(best > 1 ? (size_t)p.items * BM * HD * 4 : 0);
  p.off_pml = take(best > 1 ? (size_t)p.items * BM * 8 : 0);
  p.total = o;
  return p;
}

}  // namespace

size_t tc_xattn_bias_save_bytes(const VdetrXattnShape* s) {
  if (!s->has_bias || s->kv_heads != 1) return 0;
  const FwdPlan pl = make_plan(s);
  return (size_t)s->B * pl.nQp * pl.nKp * 16;
}
size_t tc_xattn_fwd_workspace(const VdetrXattnShape* s) { return make_plan(s).total; }

int tc_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                 const float* ref, const float* ang, const float* tables, float* out, float* lse, float* bias_save,
                 float drop_p, const unsigned long long* drop_seed, void* ws, size_t ws_bytes, cudaStream_t st) {
  const bool mqa = s->kv_heads == 1;
  if (s->has_bias && !mqa) return VDETR_ERR_UNSUPPORTED;
  if (!(drop_p >= 0.f) || drop_p >= 1.f || (drop_p > 0.f && !drop_seed)) return VDETR_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) return VDETR_ERR_BAD_ARG;   // float4 loads
  const FwdPlan pl = make_plan(s);
  if (!ws || ws_bytes < pl.total) return VDETR_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return VDETR_ERR_WORKSPACE;      // cudaMalloc / torch give >= 256 B
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);

  VdetrPack pk = {};
  pk.B = s->B; pk.nQ = s->nQ; pk.nK = s->nK; pk.nQp = pl.nQp; pk.nKp = pl.nKp; pk.kvh = s->kv_heads; pk.has_bias = s->has_bias;
  pk.q = q; pk.k = k; pk.v = v; pk.xyz = xyz; pk.ref = ref; pk.ang = (s->has_bias && s->rotate) ? ang : nullptr;
  pk.qp = reinterpret_cast<__half*>(w + pl.off_qp);
  pk.qpl = reinterpret_cast<__half*>(w + pl.off_qpl);
  pk.kp = reinterpret_cast<__half*>(w + pl.off_kp);
  pk.kpl = reinterpret_cast<__half*>(w + pl.off_kpl);
  pk.vtp = reinterpret_cast<__half*>(w + pl.off_vtp);
  pk.xyz4 = reinterpret_cast<float4*>(w + pl.off_xyz);
  pk.geo = reinterpret_cast<float4*>(w + pl.off_geo);
  vdetr_pack_kernel<<<vdetr_num_sms() * 4, 256, 0, st>>>(pk);
  VDETR_LAUNCH_CHECK();
  vdetr_pack_vt_kernel<<<s->B * s->kv_heads * (pl.nKp / 64), 256, 0, st>>>(pk);
  VDETR_LAUNCH_CHECK();

  CUtensorMap tmQ, tmQl, tmK, tmKl, tmVt;
  int rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmQ, pk.qp, (uint64_t)s->B * pl.nQp * 4, BM, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmQl, pk.qpl, (uint64_t)s->B * pl.nQp * 4, BM, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmK, pk.kp, (uint64_t)s->B * s->kv_heads * pl.nKp, BN, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmKl, pk.kpl, (uint64_t)s->B * s->kv_heads * pl.nKp, BN, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_2d(&tmVt, pk.vtp, (uint64_t)s->B * s->kv_heads * HD, (uint64_t)pl.nKp, HD, true))) return rc;

  FwdParams P = {};
  P.B = s->B; P.nQ = s->nQ; P.nK = s->nK; P.nQp = pl.nQp; P.nKp = pl.nKp; P.kvh = s->kv_heads;
  P.mtiles = pl.mtiles; P.splits = pl.splits; P.tiles_per_split = pl.tiles_per_split; P.items = pl.items;
  P.grid_n = s->has_bias ? s->grid_n : 0;
  P.log_scale = s->log_scale;
  P.c1 = s->has_bias ? (float)s->grid_n / (2.0f * 3.0f * s->max_value) : 0.f;
  P.c0 = s->has_bias ? 0.5f * (float)(s->grid_n - 1) : 0.f;
  P.xyz4 = pk.xyz4; P.geo = pk.geo; P.tables = reinterpret_cast<const float4*>(tables);
  P.out = out; P.lse = lse;
  P.bias_out = s->has_bias ? reinterpret_cast<float4*>(bias_save) : nullptr;
  P.drop_seed = drop_p > 0.f ? drop_seed : nullptr;
  P.drop_thresh = philox::thresh_of(drop_p);
  P.drop_inv_keep = 1.0f / (1.0f - drop_p);
  P.part_o = reinterpret_cast<float*>(w + pl.off_po);
  P.part_ml = reinterpret_cast<float2*>(w + pl.off_pml);

  const int table_bytes = s->has_bias ? rpe::pair_table_bytes(s->grid_n) : 0;
  const SmemLayout L = smem_layout(table_bytes);
  if (L.total + 1024 > 232448) return VDETR_ERR_UNSUPPORTED;
  const size_t smem = L.total + 1024;     // slack for the manual 1024-B alignment
  const int grid = pl.items < vdetr_num_sms() ? pl.items : vdetr_num_sms();
  VdetrTimingScope timing(s->has_bias ? VDETR_T_FWD : VDETR_T_COUNT, st);
  if (s->has_bias) {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpe_xattn_fwd_kernel<true, true><<<grid, NTHREADS, smem, st>>>(tmQ, tmQl, tmK, tmKl, tmVt, P);
  } else if (mqa) {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpe_xattn_fwd_kernel<false, true><<<grid, NTHREADS, smem, st>>>(tmQ, tmQl, tmK, tmKl, tmVt, P);
  } else {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpe_xattn_fwd_kernel<false, false><<<grid, NTHREADS, smem, st>>>(tmQ, tmQl, tmK, tmKl, tmVt, P);
  }
  VDETR_LAUNCH_CHECK();
  if (pl.splits > 1) {
    const size_t total = (size_t)s->B * pl.mtiles * BM * 4;
    int blocks = (int)((total + 255) / 256);
    fwd_combine_kernel<<<blocks, 256, 0, st>>>(P, mqa ? 1 : 0);
    VDETR_LAUNCH_CHECK();
  }
  return 0;
}
