// Vertex-RPE attention backward for sm_100a (impl = 0).  No library GEMM: every contraction is a tcgen05 MMA issued here.
//
// Pass 1 (rpe_xattn_bwd_kernel; tcgen05 + TMA, tiling / warp roles of the forward; work item = (scene, 128-row
// tile, key split)):
//     S  = Qh Kh^T + Ql Kh^T + Qh Kl^T + bias   (bit-identical to the forward; the bias is the per-pair bias the forward
//                                                saved, streamed back with bulk copies)
//     P  = exp(S - LSE)        dP = (g dO) V^T        D = rowsum(dO * O)
//     dropout (same Philox mask as the forward):  Pd = P * keep / (1 - p),   dPm = dP * keep / (1 - p)
//     g dS = P * (dPm - g D)
//     g dQ += (g dS) Kh        accumulated in TMEM over the key tiles of the item (A = the dS tile the compute warps leave
//                              in shared memory, B = the K tile exactly as TMA loaded it, read as an MN-major operand)
//   written to global memory: Pd (fp16) and g dS (fp16), both [rows][nKp]; dQ partial sums per key split.
//   g is the per-call power-of-two gradient scale (rpe_internal.h).
// Pass 2 (bwd_dkdv_kernel): dK = (g dS)^T Qh, dV = Pd^T (g dO): a TMA-fed tcgen05 GEMM whose operands are all MN-major
//   (the contraction runs over the attention rows, which is the slow index of every operand as stored).
// Pass 3: dTables from the same g dS (rpe_dtables.cu).
#include "rpe_internal.h"
#include "tc_common.cuh"
#include "philox.cuh"

namespace {

using namespace tc;

constexpr int NCOMPUTE_WARPS = 16;
constexpr int NCOMPUTE = NCOMPUTE_WARPS * 32;
constexpr int NTHREADS = NCOMPUTE + 32;
constexpr int BM = 128, BN = 64, HD = 64, QT = 32, BIAS_STRIDE_F4 = BN + 1;
constexpr float LOG2E = 1.4426950408889634f;
#ifndef VDETR_BWD_EVENT_LOOP
#define VDETR_BWD_EVENT_LOOP 0      // measured: the polling loop costs more than the reordering buys (0.41 vs 0.36 ms)
#endif
constexpr int TMEM_COLS = 512;                     // S0 [0,64) S1 [64,128) dP0 [128,192) dP1 [192,256) dQ [256,320)

struct BwdParams {
  int B, nQ, nK, nQp, nKp, kvh;
  int mtiles, splits, tiles_per_split, items;
  size_t rows_total;           // B * nQp * 4
  const float* out;            // [B,nQ,4,64] forward output
  const float* dout;           // [B,nQ,4,64]
  const float* lse;            // [B,4,nQ]
  const unsigned* absmax_bits; // bits of max|dout| -> gradient scale
  __half* pb;                  // [rows][nKp]  Pd
  __half* dsb;                 // [rows][nKp]  g * dS
  float* dq_part;              // [splits][rows][64]  g * dQ partial sums
  const float4* bias_in;       // [B][nQp][nKp] bias saved by the forward, or null (no bias)
  const unsigned long long* drop_seed;
  uint32_t drop_thresh;
  float drop_inv_keep;
};

struct SmemLayout {
  uint32_t q, ql, dO, stage, ds, bias, rowbuf, bars, total;
};
constexpr uint32_t STAGE_BYTES = 3 * BN * 128;     // K hi, K lo, V
constexpr int NSTAGE = 3;                          // K / V ring: a tile's K / V is requested two tiles ahead
__host__ __device__ inline SmemLayout smem_layout(bool has_bias) {
  SmemLayout L;
  uint32_t o = 0;
  L.q = o;      o += BM * 128;
  L.ql = o;     o += BM * 128;
  L.dO = o;     o += BM * 128;
  L.stage = o;  o += NSTAGE * STAGE_BYTES;
  L.ds = o;     o += 2 * BM * 128;
  L.bias = o;   o += has_bias ? 2 * QT * BIAS_STRIDE_F4 * 16 : 0;
  L.rowbuf = o; o += BM * 4 * 4;
  L.bars = o;   o += 256;
  L.total = o;
  return L;
}

template <bool HAS_BIAS, bool MQA>
__global__ void __launch_bounds__(NTHREADS, 1)
rpe_xattn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQl,
                     const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmKl, const __grid_constant__ CUtensorMap tmV, const BwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemLayout L = smem_layout(HAS_BIAS);
  uint8_t* sQ = smem + L.q;
  uint8_t* sQl = smem + L.ql;
  uint8_t* sdO = smem + L.dO;
  uint8_t* sStage = smem + L.stage;                 // [2] x { K hi, K lo, V }
  uint8_t* sdS = smem + L.ds;                       // [2] x [128 rows][64 keys] fp16, K-major, 128-B swizzle
  float4* sBias = reinterpret_cast<float4*>(smem + L.bias);
  float* sRow = reinterpret_cast<float*>(smem + L.rowbuf);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_full = bars + 17;   // [NSTAGE] K / V of a tile have landed in stage gj % NSTAGE
  uint64_t* bar_s = bars + 3;       // [2] S and dP of a tile are in TMEM buffer s
  uint64_t* bar_tfree = bars + 5;   // [2] every compute thread has read TMEM buffer s
  uint64_t* bar_ds = bars + 7;      // [2] the dS tile is in sdS[s]
  uint64_t* bar_dq = bars + 9;      // [2] the dQ MMAs of a tile are done: sdS[s] and the tile's K / V stage are free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint64_t* bar_bfull = bars + 13;  // [2] the bias tile of a key tile has landed in bias buffer s
  uint64_t* bar_bfree = bars + 15;  // [2] every compute thread has copied its bias values out of buffer s
  uint64_t* bar_kfree = bars + 20;  // [NSTAGE] the dQ MMAs that read K / V stage gj % NSTAGE are done (same event as bar_dq, but a
                                    // ring of NSTAGE: a parity wait may only ask for the LATEST completed phase of a barrier)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_control = warp == NCOMPUTE_WARPS;

  if (is_control) {
    if (lane == 0) {
      mbar_init(bar_q, 1);
      for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_kfree + s, 1); }
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_s + s, 1); mbar_init(bar_tfree + s, NCOMPUTE);
        mbar_init(bar_ds + s, NCOMPUTE); mbar_init(bar_dq + s, 1);
        mbar_init(bar_bfull + s, 1); mbar_init(bar_bfree + s, NCOMPUTE);
      }
      fence_barrier_init();
      prefetch_tmap(&tmQ); prefetch_tmap(&tmQl); prefetch_tmap(&tmdO); prefetch_tmap(&tmK); prefetch_tmap(&tmKl); prefetch_tmap(&tmV);
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tdP0 = tmem_base + 128, tdQ = tmem_base + 256;
  const uint32_t idesc_s = umma_idesc_f16(BM, BN);                          // S, dP: both operands K-major
  const uint32_t idesc_q = umma_idesc_f16_major(BM, HD, false, true);      // dQ = dS K: B = K tile [keys][d], MN-major
  const float gscale = vdetr_grad_scale(*P.absmax_bits);
  philox::Dropout drop;
  drop.thresh = P.drop_seed ? P.drop_thresh : 0u;
  drop.inv_keep = P.drop_inv_keep;
  {
    const unsigned long long seed = P.drop_seed ? __ldg(P.drop_seed) : 0ull;
    drop.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  }

  uint32_t g = 0, it = 0;
  for (int item = blockIdx.x; item < P.items; item += gridDim.x, ++it) {
    const int split = item % P.splits;
    const int mt = (item / P.splits) % P.mtiles;
    const int b = item / (P.splits * P.mtiles);
    const int tile_begin = split * P.tiles_per_split;
    const int T = min(P.tiles_per_split, P.nKp / BN - tile_begin);
    int q0, hsel;
    if (MQA) { q0 = mt * QT; hsel = 0; } else { q0 = (mt >> 2) * BM; hsel = mt & 3; }
    const int hk = MQA ? 0 : hsel;
    const int qrow0 = MQA ? (b * P.nQp + q0) * 4 : ((b * 4 + hsel) * P.nQp + q0);
    const int krow0 = (b * P.kvh + hk) * P.nKp;

    if (is_control) {
      if (lane == 0) {
        // (all MMAs of the previous item have completed: its last bar_dq was waited on below)
        mbar_arrive_expect_tx(bar_q, 3 * BM * 128);
        tma_load_2d(sQ, &tmQ, 0, qrow0, bar_q);
        tma_load_2d(sQl, &tmQl, 0, qrow0, bar_q);
        tma_load_2d(sdO, &tmdO, 0, qrow0, bar_q);
        auto load_kv = [&](int j) {
          const uint32_t gj = g + j;
          const int s = gj % NSTAGE;
          // stage s was last used by tile gj - NSTAGE: its dQ MMAs (which read K hi) must have completed
          if (gj >= NSTAGE) mbar_wait(bar_kfree + s, (gj / NSTAGE - 1) & 1);
          uint8_t* st = sStage + s * STAGE_BYTES;
          uint64_t* bk = bar_full + s;
          mbar_arrive_expect_tx(bk, STAGE_BYTES);
          const int kr = krow0 + (tile_begin + j) * BN;
          tma_load_2d(st, &tmK, 0, kr, bk);
          tma_load_2d(st + BN * 128, &tmKl, 0, kr, bk);
          tma_load_2d(st + 2 * BN * 128, &tmV, 0, kr, bk);
        };
        // The bias tile (32 KB, more than half of a tile's bytes) has its own ring: the compute threads copy their 16 values
        // into registers as soon as the tile has landed and release the buffer at the START of the tile's compute phase, so
        // the bias of tile j + 2 is requested a whole tile earlier than the K / V stage allows -- two bias tiles in flight.
        auto load_bias = [&](int j) {
          if (!HAS_BIAS) return;
          const uint32_t gj = g + j;
          const int s = gj & 1;
          if (gj >= 2) mbar_wait(bar_bfree + s, ((gj - 2) >> 1) & 1);
          uint64_t* bk = bar_bfull + s;
          mbar_arrive_expect_tx(bk, QT * BN * 16);
          const float4* src = P.bias_in + ((size_t)b * P.nQp + q0) * P.nKp + (tile_begin + j) * BN;
          float4* dstb = sBias + s * (QT * BIAS_STRIDE_F4);
          for (int qq = 0; qq < QT; ++qq) bulk_load_1d(dstb + qq * BIAS_STRIDE_F4, src + (size_t)qq * P.nKp, BN * 16, bk);
        };
        auto issue_dq = [&](int j) {
          const uint32_t gj = g + j;
          const int s = gj & 1;
          mbar_wait_relaxed(bar_ds + s, (gj >> 1) & 1);      // a tile of compute work away
          tc_fence_after();
          const uint32_t a0 = smem_u32(sdS + s * (BM * 128)), b0 = smem_u32(sStage + (gj % NSTAGE) * STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_bf16(tdQ, umma_desc_sw128(a0) + (uint64_t)(kk * 2), umma_desc_sw128_mn(b0 + kk * 2048, 8192, 1024), idesc_q,
                      (j > 0) || (kk > 0));
          umma_commit(bar_dq + s);
          umma_commit(bar_kfree + gj % NSTAGE);
        };
        auto issue_s = [&](int j) {
          const uint32_t gj = g + j;
          const int s = gj & 1;
          mbar_wait(bar_full + gj % NSTAGE, (gj / NSTAGE) & 1);
          if (gj >= 2) mbar_wait(bar_tfree + s, ((gj - 2) >> 1) & 1);      // TMEM buffer s has been read
          tc_fence_after();
          {
            const uint8_t* st = sStage + (gj % NSTAGE) * STAGE_BYTES;
            const uint64_t dqh = umma_desc_sw128(smem_u32(sQ)), dql = umma_desc_sw128(smem_u32(sQl));
            const uint64_t dkh = umma_desc_sw128(smem_u32(st)), dkl = umma_desc_sw128(smem_u32(st + BN * 128));
            const uint64_t ddo = umma_desc_sw128(smem_u32(sdO)), dv = umma_desc_sw128(smem_u32(st + 2 * BN * 128));
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) umma_bf16(tS0 + s * BN, dqh + (uint64_t)(kk * 2), dkh + (uint64_t)(kk * 2), idesc_s, kk > 0);
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) umma_bf16(tS0 + s * BN, dql + (uint64_t)(kk * 2), dkh + (uint64_t)(kk * 2), idesc_s, true);
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) umma_bf16(tS0 + s * BN, dqh + (uint64_t)(kk * 2), dkl + (uint64_t)(kk * 2), idesc_s, true);
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) umma_bf16(tdP0 + s * BN, ddo + (uint64_t)(kk * 2), dv + (uint64_t)(kk * 2), idesc_s, kk > 0);
            umma_commit(bar_s + s);
          }
        };
#if VDETR_BWD_EVENT_LOOP
        // The control lane serves four queues (bias loads, K / V loads, S / dP MMAs, dQ MMAs); each entry has its own
        // readiness condition on an mbarrier.  Polling them round-robin issues whatever is ready first: a late K / V tile no
        // longer delays the dQ MMAs of the tile before it (which the compute warps of the NEXT tile wait for), and loads go
        // out the moment their buffer is released.
        {
          auto ready = [&](uint64_t* bar, uint32_t gj_done) { return mbar_try_wait(bar + (gj_done & 1), (gj_done >> 1) & 1); };
          int nb = 0, nk = 0, ns = 0, nd = 0;
          load_kv(0); load_bias(0); nb = nk = 1;
          mbar_wait(bar_q, it & 1);
          while (nd < T) {
            bool progress = false;
            if (HAS_BIAS && nb < T && (g + nb < 2 || ready(bar_bfree, g + nb - 2))) { load_bias(nb++); progress = true; }
            if (nk < T && (g + nk < NSTAGE || mbar_try_wait(bar_kfree + (g + nk) % NSTAGE, ((g + nk) / NSTAGE - 1) & 1))) { load_kv(nk++); progress = true; }
            if (ns < nk && (!HAS_BIAS || ns < nb) && mbar_try_wait(bar_full + (g + ns) % NSTAGE, ((g + ns) / NSTAGE) & 1) &&
                (g + ns < 2 || ready(bar_tfree, g + ns - 2))) {
              issue_s(ns++); progress = true;
            }
            if (nd < ns && ready(bar_ds, g + nd)) { issue_dq(nd++); progress = true; }
            if (!progress) __nanosleep(64);
          }
        }
#else
        load_kv(0); load_bias(0);
        if (T > 1) { load_kv(1); load_bias(1); }
        if (T > 2) load_kv(2);
        mbar_wait(bar_q, it & 1);
        for (int j = 0; j < T; ++j) {
          if (j > 0 && j + 1 < T) load_bias(j + 1);           // buffer of tile j - 1: released when its compute phase began
          issue_s(j);
          if (j > 0) {
            issue_dq(j - 1);
            if (j + 2 < T) load_kv(j + 2);                  // into the stage tile j - 1 just released
          }
        }
        issue_dq(T - 1);
#endif
        // sQ / sQl / sdO / the stages may only be overwritten (next item) after the last MMAs have finished
        mbar_wait(bar_dq + ((g + T - 1) & 1), ((g + T - 1) >> 1) & 1);
      }
      __syncwarp();
    } else {
      const int quarter = warp & 3, slice = warp >> 2;
      const int row = quarter * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      const int q = MQA ? q0 + (row >> 2) : q0 + row;
      const int h = MQA ? (row & 3) : hsel;
      // D = rowsum(dO * O): each thread covers 16 of the 64 columns of its row
      float dpart = 0.f;
      if (q < P.nQ) {
        const float4* o4 = reinterpret_cast<const float4*>(P.out + (((size_t)b * P.nQ + q) * 4 + h) * HD + slice * 16);
        const float4* d4 = reinterpret_cast<const float4*>(P.dout + (((size_t)b * P.nQ + q) * 4 + h) * HD + slice * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 a = __ldg(o4 + c), d = __ldg(d4 + c);
          dpart += a.x * d.x + a.y * d.y + a.z * d.z + a.w * d.w;
        }
      }
      sRow[row * 4 + slice] = dpart;
      named_bar_sync(1, NCOMPUTE);
      const float4 dp4 = *reinterpret_cast<const float4*>(sRow + row * 4);
      const float Drow = ((dp4.x + dp4.y) + (dp4.z + dp4.w)) * gscale;
      const float lse2 = (q < P.nQ) ? __ldg(P.lse + ((size_t)b * 4 + h) * P.nQ + q) * LOG2E : INFINITY;
      const size_t grow = (size_t)(qrow0 + row) * P.nKp;

      uint32_t pk[8], dk_[8];                                               // Pd / g dS of the previous tile, stored one tile late
      int key0_prev = -1;
      auto store_prev = [&]() {
        if (key0_prev < 0) return;
        uint4* dstp = reinterpret_cast<uint4*>(P.pb + grow + key0_prev + slice * 16);
        uint4* dstd = reinterpret_cast<uint4*>(P.dsb + grow + key0_prev + slice * 16);
        dstp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]); dstp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        dstd[0] = make_uint4(dk_[0], dk_[1], dk_[2], dk_[3]); dstd[1] = make_uint4(dk_[4], dk_[5], dk_[6], dk_[7]);
      };
      for (int j = 0; j < T; ++j) {
        const uint32_t gj = g + j;
        const int s = gj & 1;
        const int key0 = (tile_begin + j) * BN;
        float bz[16];
        if (HAS_BIAS) {
          mbar_wait(bar_bfull + s, (gj >> 1) & 1);                          // the bias tile has landed
          const float* brow =
              reinterpret_cast<const float*>(sBias + s * (QT * BIAS_STRIDE_F4) + (row >> 2) * BIAS_STRIDE_F4 + slice * 16) + (row & 3);
#pragma unroll
          for (int c = 0; c < 16; ++c) bz[c] = brow[c * 4];
          // generic-proxy reads of the buffer are ordered before the bulk copies (async proxy) that refill it
          fence_proxy_async_smem();
          mbar_arrive(bar_bfree + s);
        }
        // Pd and g dS of the previous tile go to global memory here: a proxy fence waits for the thread's outstanding stores,
        // so stores issued right before one sat on the tile's critical path (12 % of the stall samples); issued here they
        // drain during this tile's compute phase.
        store_prev();
        mbar_wait(bar_s + s, (gj >> 1) & 1);
        tc_fence_after();
        uint32_t sr[16], dr[16];
        tmem_ld16(tS0 + s * BN + lane_addr + slice * 16, sr);
        tmem_ld16(tdP0 + s * BN + lane_addr + slice * 16, dr);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_tfree + s);                                         // TMEM buffer s may be overwritten
        const bool tail_tile = key0 + BN > P.nK;
        uint32_t keep = 0xFFFFu;
        float dscale = 1.f;
        if (drop.thresh) {
          keep = philox::keep_mask16(drop, (uint32_t)(qrow0 + row), (uint32_t)((key0 >> 4) + slice));
          dscale = drop.inv_keep;
        }
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float sv = __uint_as_float(sr[c + e]);
            if (HAS_BIAS) sv += bz[c + e];
            float p = ex2_approx(sv * LOG2E - lse2);
            if (tail_tile && key0 + slice * 16 + c + e >= P.nK) p = 0.f;       // only the last key tile is ragged
            const float m = ((keep >> (c + e)) & 1u) ? dscale : 0.f;
            const float ds = p * (__uint_as_float(dr[c + e]) * m - Drow);      // = g * dS
            pv[e] = p * m; dv[e] = ds;
          }
          pk[c >> 1] = pack_f16x2(pv[0], pv[1]);
          dk_[c >> 1] = pack_f16x2(dv[0], dv[1]);
        }
        // dS tile -> sdS[s] as the A operand of dQ += dS K (K-major rows of 128 B, 128-byte swizzle: chunk ^= row % 8)
        if (gj >= 2) mbar_wait(bar_dq + s, ((gj - 2) >> 1) & 1);             // the dQ MMAs of tile gj - 2 have read sdS[s]
        {
          uint8_t* prow = sdS + s * (BM * 128) + (row >> 3) * 1024 + (row & 7) * 128;
          const int ch0 = slice * 2, ch1 = slice * 2 + 1;
          *reinterpret_cast<uint4*>(prow + ((ch0 ^ (row & 7)) << 4)) = make_uint4(dk_[0], dk_[1], dk_[2], dk_[3]);
          *reinterpret_cast<uint4*>(prow + ((ch1 ^ (row & 7)) << 4)) = make_uint4(dk_[4], dk_[5], dk_[6], dk_[7]);
        }
        // generic-proxy writes of sdS are ordered before the async-proxy reads that follow the arrive (the dQ MMAs)
        fence_proxy_async_smem();
        mbar_arrive(bar_ds + s);
        key0_prev = key0;
      }
      store_prev();
      // ------------------------------------------------------------------ item epilogue: dQ partial sums of this split
      {
        const uint32_t gl = g + T - 1;
        mbar_wait(bar_dq + (gl & 1), (gl >> 1) & 1);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(tdQ + lane_addr + slice * 16, o);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(P.dq_part + ((size_t)split * P.rows_total + (size_t)(qrow0 + row)) * HD + slice * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_float4(__uint_as_float(o[4 * c]), __uint_as_float(o[4 * c + 1]), __uint_as_float(o[4 * c + 2]),
                               __uint_as_float(o[4 * c + 3]));
        tc_fence_before();
      }
      named_bar_sync(2, NCOMPUTE);                         // sRow reusable by the next item
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

// ------------------------------------------------------------------------------------------------ pass 2: dK, dV
// Work item = (batch entry e = scene (shared K/V) or scene x head, block of 128 keys).  Both outputs contract over the
// attention rows of e:   dK[key][d] = sum_r dS[r][key] Qh[r][d],   dV[key][d] = sum_r Pd[r][key] dO[r][d].
// Per 64-row step one TMA stage = dS [64 r x 128 keys] + Pd [64 r x 128 keys] + Qh [64 r x 64] + dO [64 r x 64] (48 KB),
// all consumed as MN-major operands (rows are the K dimension of the MMA); accumulators: 2 x [128 keys x 64] fp32 in
// TMEM, double buffered so that the epilogue of an item overlaps the main loop of the next one.
constexpr int KV_BLOCK = 128, KV_ROWS = 64, KV_STAGES = 4;
constexpr uint32_t KV_STAGE_BYTES = 2 * KV_ROWS * KV_BLOCK * 2 + 2 * KV_ROWS * HD * 2;      // 49152
constexpr int KV_THREADS = 192;                     // warps 0-3: epilogue (TMEM quarter = warp), 4: TMA, 5: MMA

struct DkdvParams {
  int E, kvh, nK, nKp, rows_e, kblocks, items;
  float* dk;                    // [B][nK][kvh][64]
  float* dv;
  const unsigned* absmax_bits;
};

__global__ void __launch_bounds__(KV_THREADS, 1)
bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmP,
                const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO, const DkdvParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KV_STAGES * KV_STAGE_BYTES);
  uint64_t* bar_full = bars;                    // [KV_STAGES]
  uint64_t* bar_empty = bars + KV_STAGES;       // [KV_STAGES]
  uint64_t* bar_accf = bars + 2 * KV_STAGES;    // [2] accumulators of an item complete
  uint64_t* bar_acce = bars + 2 * KV_STAGES + 2;  // [2] accumulators drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * KV_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < KV_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(bar_accf + a, 1); mbar_init(bar_acce + a, 128); }
      fence_barrier_init();
      prefetch_tmap(&tmDS); prefetch_tmap(&tmP); prefetch_tmap(&tmQ); prefetch_tmap(&tmdO);
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ksteps = P.rows_e / KV_ROWS;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
        const int kb = item % P.kblocks, e = item / P.kblocks;
        const int key0 = kb * KV_BLOCK, row0 = e * P.rows_e;
        for (int ks = 0; ks < ksteps; ++ks, ++n) {
          const int s = n % KV_STAGES;
          if (n >= KV_STAGES) mbar_wait(bar_empty + s, ((n / KV_STAGES) - 1) & 1);
          uint8_t* st = smem + s * KV_STAGE_BYTES;
          mbar_arrive_expect_tx(bar_full + s, KV_STAGE_BYTES);
          const int r = row0 + ks * KV_ROWS;
          tma_load_2d(st, &tmDS, key0, r, bar_full + s);                     // dS keys [key0, +64)   (out-of-range keys: zeros)
          tma_load_2d(st + 8192, &tmDS, key0 + 64, r, bar_full + s);         // dS keys [key0+64, +64)
          tma_load_2d(st + 16384, &tmP, key0, r, bar_full + s);
          tma_load_2d(st + 24576, &tmP, key0 + 64, r, bar_full + s);
          tma_load_2d(st + 32768, &tmQ, 0, r, bar_full + s);
          tma_load_2d(st + 40960, &tmdO, 0, r, bar_full + s);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16_major(KV_BLOCK, HD, true, true);
      uint32_t n = 0, il = 0;
      for (int item = blockIdx.x; item < P.items; item += gridDim.x, ++il) {
        const int a = il & 1;
        if (il >= 2) mbar_wait(bar_acce + a, ((il >> 1) - 1) & 1);           // epilogue drained this accumulator pair
        tc_fence_after();
        const uint32_t tK = tmem_base + a * 128, tV = tK + 64;
        for (int ks = 0; ks < ksteps; ++ks, ++n) {
          const int s = n % KV_STAGES;
          mbar_wait(bar_full + s, (n / KV_STAGES) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * KV_STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < KV_ROWS / 16; ++kk) {
            // 16 rows per MMA = two 8-row groups (1024 B apart); the two 64-key blocks of A are 8192 B apart
            umma_bf16(tK, umma_desc_sw128_mn(st + kk * 2048, 8192, 1024), umma_desc_sw128_mn(st + 32768 + kk * 2048, 8192, 1024),
                      idesc, (ks > 0) || (kk > 0));
            umma_bf16(tV, umma_desc_sw128_mn(st + 16384 + kk * 2048, 8192, 1024),
                      umma_desc_sw128_mn(st + 40960 + kk * 2048, 8192, 1024), idesc, (ks > 0) || (kk > 0));
          }
          umma_commit(bar_empty + s);
        }
        umma_commit(bar_accf + a);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> dk / dv (user layout, un-scaled)
    const float ginv = 1.0f / vdetr_grad_scale(*P.absmax_bits);
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    uint32_t il = 0;
    for (int item = blockIdx.x; item < P.items; item += gridDim.x, ++il) {
      const int a = il & 1;
      const int kb = item % P.kblocks, e = item / P.kblocks;
      const int key = kb * KV_BLOCK + warp * 32 + lane;
      const int bsc = e / P.kvh, hk = e % P.kvh;
      mbar_wait(bar_accf + a, (il >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        float* base = (t == 0 ? P.dk : P.dv) + (((size_t)bsc * P.nK + key) * P.kvh + hk) * HD;
#pragma unroll 1
        for (int c0 = 0; c0 < HD; c0 += 16) {
          uint32_t o[16];
          tmem_ld16(tmem_base + a * 128 + t * 64 + c0 + lane_addr, o);
          tmem_ld_wait();
          if (key < P.nK) {
            float4* dst = reinterpret_cast<float4*>(base + c0);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              dst[c] = make_float4(__uint_as_float(o[4 * c]) * ginv, __uint_as_float(o[4 * c + 1]) * ginv,
                                   __uint_as_float(o[4 * c + 2]) * ginv, __uint_as_float(o[4 * c + 3]) * ginv);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acce + a);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// dQ partial sums (per key split, padded packed rows, scaled by g) -> dq in the user layout
struct UnpackParams {
  int B, nQ, nQp, kvh, splits;
  size_t rows_total;
  const float* dq_part;
  float* dq;
  const unsigned* absmax_bits;
};
__global__ void bwd_unpack_dq_kernel(UnpackParams U) {
  const float ginv = 1.0f / vdetr_grad_scale(*U.absmax_bits);
  const size_t n4 = (size_t)U.B * U.nQ * 4 * 16;                     // float4 elements of dq
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int d4 = (int)(i & 15), h = (int)((i >> 4) & 3);
    const size_t bq = i >> 6;
    const int q = (int)(bq % U.nQ), b = (int)(bq / U.nQ);
    const size_t row = U.kvh == 1 ? ((size_t)b * U.nQp + q) * 4 + h : ((size_t)b * 4 + h) * U.nQp + q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < U.splits; ++s) {                             // fixed order: dq is run-to-run deterministic
      const float4 v = __ldg(reinterpret_cast<const float4*>(U.dq_part + ((size_t)s * U.rows_total + row) * 64) + d4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(U.dq)[i] = make_float4(acc.x * ginv, acc.y * ginv, acc.z * ginv, acc.w * ginv);
  }
}

struct BwdPlan {
  int nQp, nKp, mtiles, splits, tiles_per_split, items;
  size_t off_qp, off_qpl, off_dop, off_kp, off_kpl, off_vp, off_max, off_xyz, off_geo, off_pb, off_dsb, off_dt, off_dqp,
      off_bias, off_fwd, off_out, total;
};
BwdPlan make_plan(const VdetrXattnShape* s, bool recompute_bias) {
  BwdPlan p;
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
    if (sp > 1 && ktiles / sp < 4) break;
    const int tps = (ktiles + sp - 1) / sp;
    if ((ktiles + tps - 1) / tps != sp) continue;
    const double waves = (double)(base * sp) / sms;
    const double eff = waves / (double)((long)((base * sp + sms - 1) / sms));
    if (eff > best_eff + 0.04) { best_eff = eff; best = sp; }
    if (eff >= 0.93) { best = sp; break; }
  }
  p.splits = best;
  p.tiles_per_split = (ktiles + best - 1) / best;
  p.items = (int)(base * best);
  const size_t rows = (size_t)s->B * p.nQp * 4;
  const size_t krows = (size_t)s->B * s->kv_heads * p.nKp;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += vdetr_align_up(bytes, 1024); return r; };
  p.off_qp = take(rows * 64 * 2);
  p.off_qpl = take(rows * 64 * 2);
  p.off_dop = take(rows * 64 * 2);
  p.off_kp = take(krows * 64 * 2);
  p.off_kpl = take(krows * 64 * 2);
  p.off_vp = take(krows * 64 * 2);
  p.off_max = take(16);
  p.off_xyz = take(s->has_bias ? (size_t)s->B * p.nKp * 16 : 0);
  p.off_geo = take(s->has_bias ? (size_t)s->B * p.nQp * 9 * 16 : 0);
  p.off_pb = take(rows * p.nKp * 2);
  p.off_dsb = take(rows * p.nKp * 2);
  p.off_dt = take(s->has_bias ? rpe_dtables_scratch_bytes(s) : 0);
  p.off_dqp = take((size_t)best * rows * 64 * 4);
  // no saved bias: the forward kernel is run once more into a transient buffer (same bits as the forward pass produced)
  const bool rb = s->has_bias && recompute_bias;
  p.off_bias = take(rb ? tc_xattn_bias_save_bytes(s) : 0);
  p.off_fwd = take(rb ? tc_xattn_fwd_workspace(s) : 0);
  p.off_out = take(rb ? (size_t)s->B * s->nQ * 4 * 64 * 4 + (size_t)s->B * 4 * s->nQ * 4 : 0);
  p.total = o;
  return p;
}

}  // namespace

size_t tc_xattn_bwd_workspace(const VdetrXattnShape* s, int bias_is_saved) { return make_plan(s, !bias_is_saved).total; }

int tc_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                 const float* ref, const float* ang, const float* tables, const float* out, const float* lse,
                 const float* dout, const float* bias_saved, float drop_p, const unsigned long long* drop_seed, float* dq,
                 float* dk, float* dv, float* dtables, void* ws, size_t ws_bytes, cudaStream_t st) {
  const bool mqa = s->kv_heads == 1;
  if (s->has_bias && !mqa) return VDETR_ERR_UNSUPPORTED;
  if (!(drop_p >= 0.f) || drop_p >= 1.f || (drop_p > 0.f && !drop_seed)) return VDETR_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dout)) & 15)
    return VDETR_ERR_BAD_ARG;                                 // the pack kernel uses float4 loads
  const bool recompute = s->has_bias && bias_saved == nullptr;
  const BwdPlan pl = make_plan(s, recompute);
  if (!ws || ws_bytes < pl.total || (reinterpret_cast<uintptr_t>(ws) & 255) != 0) return VDETR_ERR_WORKSPACE;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  int rc;

  if (recompute) {
    float* o2 = reinterpret_cast<float*>(w + pl.off_out);
    float* lse2 = o2 + (size_t)s->B * s->nQ * 4 * 64;
    float* bias_ws = reinterpret_cast<float*>(w + pl.off_bias);
    if ((rc = tc_xattn_fwd(s, q, k, v, xyz, ref, ang, tables, o2, lse2, bias_ws, 0.f, nullptr, w + pl.off_fwd,
                           tc_xattn_fwd_workspace(s), st)))
      return rc;
    bias_saved = bias_ws;
  }

  VdetrPack pk = {};
  pk.B = s->B; pk.nQ = s->nQ; pk.nK = s->nK; pk.nQp = pl.nQp; pk.nKp = pl.nKp; pk.kvh = s->kv_heads; pk.has_bias = s->has_bias;
  pk.q = q; pk.k = k; pk.v = v; pk.xyz = xyz; pk.ref = ref; pk.ang = (s->has_bias && s->rotate) ? ang : nullptr; pk.dout = dout;
  unsigned* absmax = reinterpret_cast<unsigned*>(w + pl.off_max);
  VDETR_CUDA_TRY(cudaMemsetAsync(absmax, 0, 4, st));
  vdetr_absmax_kernel<<<vdetr_num_sms() * 2, 256, 0, st>>>(dout, (size_t)s->B * s->nQ * 4 * 64, absmax);
  VDETR_LAUNCH_CHECK();
  pk.dout_absmax_bits = absmax;
  pk.qp = reinterpret_cast<__half*>(w + pl.off_qp);
  pk.qpl = reinterpret_cast<__half*>(w + pl.off_qpl);
  pk.dop = reinterpret_cast<__half*>(w + pl.off_dop);
  pk.kp = reinterpret_cast<__half*>(w + pl.off_kp);
  pk.kpl = reinterpret_cast<__half*>(w + pl.off_kpl);
  pk.vp = reinterpret_cast<__half*>(w + pl.off_vp);
  pk.vtp = nullptr;
  pk.xyz4 = reinterpret_cast<float4*>(w + pl.off_xyz);
  pk.geo = reinterpret_cast<float4*>(w + pl.off_geo);
  vdetr_pack_kernel<<<vdetr_num_sms() * 4, 256, 0, st>>>(pk);
  VDETR_LAUNCH_CHECK();

  const uint64_t rows = (uint64_t)s->B * pl.nQp * 4, krows = (uint64_t)s->B * s->kv_heads * pl.nKp;
  CUtensorMap tmQ, tmQl, tmdO, tmK, tmKl, tmV;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmQ, pk.qp, rows, BM, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmQl, pk.qpl, rows, BM, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmdO, pk.dop, rows, BM, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmK, pk.kp, krows, BN, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmKl, pk.kpl, krows, BN, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmV, pk.vp, krows, BN, true))) return rc;

  BwdParams P = {};
  P.B = s->B; P.nQ = s->nQ; P.nK = s->nK; P.nQp = pl.nQp; P.nKp = pl.nKp; P.kvh = s->kv_heads;
  P.mtiles = pl.mtiles; P.splits = pl.splits; P.tiles_per_split = pl.tiles_per_split; P.items = pl.items;
  P.rows_total = (size_t)rows;
  P.out = out; P.dout = dout; P.lse = lse;
  P.absmax_bits = absmax;
  P.pb = reinterpret_cast<__half*>(w + pl.off_pb);
  P.dsb = reinterpret_cast<__half*>(w + pl.off_dsb);
  P.dq_part = reinterpret_cast<float*>(w + pl.off_dqp);
  P.bias_in = s->has_bias ? reinterpret_cast<const float4*>(bias_saved) : nullptr;
  P.drop_seed = drop_p > 0.f ? drop_seed : nullptr;
  P.drop_thresh = philox::thresh_of(drop_p);
  P.drop_inv_keep = 1.0f / (1.0f - drop_p);

  {
    const SmemLayout L = smem_layout(s->has_bias != 0);
    const size_t smem = L.total + 1024;
    if (smem > 232448) return VDETR_ERR_UNSUPPORTED;
    const int grid = pl.items < vdetr_num_sms() ? pl.items : vdetr_num_sms();
    VdetrTimingScope timing(s->has_bias ? VDETR_T_BWD : VDETR_T_COUNT, st);
    if (s->has_bias) {
      VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rpe_xattn_bwd_kernel<true, true><<<grid, NTHREADS, smem, st>>>(tmQ, tmQl, tmdO, tmK, tmKl, tmV, P);
    } else if (mqa) {
      VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rpe_xattn_bwd_kernel<false, true><<<grid, NTHREADS, smem, st>>>(tmQ, tmQl, tmdO, tmK, tmKl, tmV, P);
    } else {
      VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rpe_xattn_bwd_kernel<false, false><<<grid, NTHREADS, smem, st>>>(tmQ, tmQl, tmdO, tmK, tmKl, tmV, P);
    }
  }
  VDETR_LAUNCH_CHECK();

  // ---- dQ: sum of the per-split partial sums, un-scaled, user layout
  {
    UnpackParams U = {s->B, s->nQ, pl.nQp, s->kv_heads, pl.splits, (size_t)rows, P.dq_part, dq, absmax};
    bwd_unpack_dq_kernel<<<vdetr_num_sms() * 4, 256, 0, st>>>(U);
    VDETR_LAUNCH_CHECK();
  }

  // ---- pass 2: dK = dS^T Q, dV = Pd^T dO
  {
    const int E = mqa ? s->B : s->B * 4;
    const int rows_e = mqa ? pl.nQp * 4 : pl.nQp;
    CUtensorMap tmDS, tmP, tmQ64, tmdO64;
    if ((rc = vdetr_make_tmap_bf16_2d(&tmDS, P.dsb, rows, (uint64_t)pl.nKp, KV_ROWS, true))) return rc;
    if ((rc = vdetr_make_tmap_bf16_2d(&tmP, P.pb, rows, (uint64_t)pl.nKp, KV_ROWS, true))) return rc;
    if ((rc = vdetr_make_tmap_bf16_rows64(&tmQ64, pk.qp, rows, KV_ROWS, true))) return rc;
    if ((rc = vdetr_make_tmap_bf16_rows64(&tmdO64, pk.dop, rows, KV_ROWS, true))) return rc;
    DkdvParams D = {};
    D.E = E; D.kvh = s->kv_heads; D.nK = s->nK; D.nKp = pl.nKp; D.rows_e = rows_e;
    D.kblocks = (pl.nKp + KV_BLOCK - 1) / KV_BLOCK;
    D.items = E * D.kblocks;
    D.dk = dk; D.dv = dv; D.absmax_bits = absmax;
    const size_t smem = (size_t)KV_STAGES * KV_STAGE_BYTES + 256 + 1024;
    const int grid = D.items < vdetr_num_sms() ? D.items : vdetr_num_sms();
    VdetrTimingScope timing(s->has_bias ? VDETR_T_BWD2 : VDETR_T_COUNT, st);
    VDETR_CUDA_TRY(cudaFuncSetAttribute(bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_dkdv_kernel<<<grid, KV_THREADS, smem, st>>>(tmDS, tmP, tmQ64, tmdO64, D);
    VDETR_LAUNCH_CHECK();
  }

  // ---- pass 3: dTables
  if (s->has_bias) {
    if ((rc = rpe_dtables_launch(s, pl.nQp, pl.nKp, pk.xyz4, pk.geo, P.dsb, absmax, 0, dtables, w + pl.off_dt,
                                 rpe_dtables_scratch_bytes(s), st)))
      return rc;
  }
  return 0;
}
