// Vertex-RPE attention backward for sm_100a (impl = 0).
//
// Pass 1 (this file's kernel, tcgen05 + TMA, same tiling / warp roles as the forward):
//     S  = Q K^T + rpe            (bias recomputed, never read from memory)
//     P  = exp(S - LSE)           dP = dO V^T          dS = P * (dP - D),   D = rowsum(dO * O)
//   written once as  P (fp16) and g*dS (fp16), both [rows][nKp];
//   g is the per-call power-of-two gradient scale (rpe_internal.h).
// Pass 2: the three plain GEMMs   dQ = dS K,  dK = dS^T Q,  dV = P^T dO   (cuBLAS, fp16 in / fp32 out).
// Pass 3: dTables from the same g*dS (rpe_dtables.cu).
#include <cublas_v2.h>
#include "rpe_internal.h"
#include "tc_common.cuh"
#include "rpe_fast.cuh"

namespace {

using namespace tc;
using rpe::rpe_bias_pair;

constexpr int NCOMPUTE_WARPS = 16;
constexpr int NCOMPUTE = NCOMPUTE_WARPS * 32;
constexpr int NTHREADS = NCOMPUTE + 32;
constexpr int BM = 128, BN = 64, HD = 64, QT = 32, GEO_F4 = 9, BIAS_STRIDE_F4 = BN + 1;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int TMEM_COLS = 256;                     // S0 [0,64) S1 [64,128) dP0 [128,192) dP1 [192,256)

struct BwdParams {
  int B, nQ, nK, nQp, nKp, kvh;
  int mtiles, splits, tiles_per_split, items;
  int grid_n;
  float log_scale, c1, c0;
  const float4* xyz4;
  const float4* geo;
  const float4* tables;
  const float* out;            // [B,nQ,4,64] forward output
  const float* dout;           // [B,nQ,4,64]
  const float* lse;            // [B,4,nQ]
  const unsigned* absmax_bits; // bits of max|dout| -> gradient scale
  __half* pb;                  // [rows][nKp]  P
  __half* dsb;                 // [rows][nKp]  g * dS
  const float4* bias_in;       // [B][nQp][nKp] bias saved by the forward (SAVED), else null
};

struct SmemLayout {
  uint32_t q, dO, k, v, tables, bias, xyz, geo, rowbuf, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int table_bytes, int bias_bufs = 1) {
  SmemLayout L;
  uint32_t o = 0;
  L.q = o;      o += BM * 128;
  L.dO = o;     o += BM * 128;
  L.k = o;      o += BN * 128;
  L.v = o;      o += BN * 128;
  L.tables = o; o += (uint32_t)((table_bytes + 1023) / 1024 * 1024);
  L.bias = o;   o += bias_bufs * QT * BIAS_STRIDE_F4 * 16;
  L.xyz = o;    o += 2 * BN * 16;
  L.geo = o;    o += QT * GEO_F4 * 16;
  L.rowbuf = o; o += BM * 4 * 4;
  L.bars = o;   o += 128;
  L.total = o;
  return L;
}

// SAVED: the bias of every pair was stored by the forward; it is streamed back with bulk copies (double buffered)
// instead of being recomputed, and the kernel needs neither tables nor query geometry.
template <bool HAS_BIAS, bool MQA, bool SAVED>
__global__ void __launch_bounds__(NTHREADS, 1)
rpe_xattn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const BwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemLayout L = smem_layout((HAS_BIAS && !SAVED) ? rpe::pair_table_bytes(P.grid_n) : 0, SAVED ? 2 : 1);
  uint8_t* sQ = smem + L.q;
  uint8_t* sdO = smem + L.dO;
  uint8_t* sK = smem + L.k;
  uint8_t* sV = smem + L.v;
  const char* sTab = reinterpret_cast<const char*>(smem + L.tables);
  float4* sBias = reinterpret_cast<float4*>(smem + L.bias);
  float4* sXyz = reinterpret_cast<float4*>(smem + L.xyz);
  float4* sGeo = reinterpret_cast<float4*>(smem + L.geo);
  float* sRow = reinterpret_cast<float*>(smem + L.rowbuf);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;       // [2]  per smem xyz / TMEM buffer parity, so that no waiter can ever be
  uint64_t* bar_s = bars + 3;       // [2]  two phases behind the barrier it waits on
  uint64_t* bar_p = bars + 5;       // [2]
  uint64_t* bar_kfree = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_control = warp == NCOMPUTE_WARPS;

  if (is_control) {
    if (lane == 0) {
      mbar_init(bar_q, 1); mbar_init(bar_k + 0, 1); mbar_init(bar_k + 1, 1); mbar_init(bar_kfree, 1);
      mbar_init(bar_s + 0, 1); mbar_init(bar_s + 1, 1); mbar_init(bar_p + 0, NCOMPUTE); mbar_init(bar_p + 1, NCOMPUTE);
      fence_barrier_init();
      prefetch_tmap(&tmQ); prefetch_tmap(&tmdO); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(tmem_slot);
  } else if (HAS_BIAS && !SAVED) {
    // tables -> shared memory as fp16 x-pairs, once per (persistent) CTA
    rpe::load_pair_tables(reinterpret_cast<uint4*>(smem + L.tables), P.tables, P.grid_n, tid, NCOMPUTE);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tdP0 = tmem_base + 128;
  const uint32_t idesc_s = umma_idesc_f16(BM, BN);     // S  = Q K^T : fp16 operands, exactly as the forward
  const uint32_t idesc_d = umma_idesc_f16(BM, BN);     // g*dP = (g dO) V^T: scaled fp16
  const float gscale = vdetr_grad_scale(*P.absmax_bits);

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
        // all MMAs of the previous item have completed (their tiles were consumed: bar_p waited below)
        mbar_arrive_expect_tx(bar_q, 2 * BM * 128);
        tma_load_2d(sQ, &tmQ, 0, qrow0, bar_q);
        tma_load_2d(sdO, &tmdO, 0, qrow0, bar_q);
        const uint32_t kbytes = 2 * BN * 128 + (HAS_BIAS ? (SAVED ? QT * BN * 16 : BN * 16) : 0);
        for (int j = 0; j < T; ++j) {
          const uint32_t gj = g + j;
          uint64_t* bk = bar_k + (gj & 1);
          // tile gj-2 used the same TMEM / xyz buffers and the same bar_k: its readers must be done
          if (gj >= 2) mbar_wait(bar_p + (gj & 1), ((gj - 2) >> 1) & 1);
          if (j > 0) mbar_wait(bar_kfree, (gj - 1) & 1);       // MMAs of tile j-1 done: sK / sV free
          mbar_arrive_expect_tx(bk, kbytes);
          tma_load_2d(sK, &tmK, 0, krow0 + (tile_begin + j) * BN, bk);
          tma_load_2d(sV, &tmV, 0, krow0 + (tile_begin + j) * BN, bk);
          if (HAS_BIAS && !SAVED) bulk_load_1d(sXyz + (gj & 1) * BN, P.xyz4 + (size_t)b * P.nKp + (tile_begin + j) * BN, BN * 16, bk);
          if (HAS_BIAS && SAVED) {
            const float4* src = P.bias_in + ((size_t)b * P.nQp + q0) * P.nKp + (tile_begin + j) * BN;
            float4* dstb = sBias + (gj & 1) * (QT * BIAS_STRIDE_F4);
            for (int qq = 0; qq < QT; ++qq) bulk_load_1d(dstb + qq * BIAS_STRIDE_F4, src + (size_t)qq * P.nKp, BN * 16, bk);
          }
          if (j == 0) mbar_wait(bar_q, it & 1);
          mbar_wait(bk, (gj >> 1) & 1);
          tc_fence_after();
          const uint64_t dq = umma_desc_sw128(smem_u32(sQ)), dk = umma_desc_sw128(smem_u32(sK));
          const uint64_t ddo = umma_desc_sw128(smem_u32(sdO)), dv = umma_desc_sw128(smem_u32(sV));
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tS0 + (gj & 1) * BN, dq + (uint64_t)(kk * 2), dk + (uint64_t)(kk * 2), idesc_s, kk > 0);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tdP0 + (gj & 1) * BN, ddo + (uint64_t)(kk * 2), dv + (uint64_t)(kk * 2), idesc_d, kk > 0);
          umma_commit(bar_s + (gj & 1));
          umma_commit(bar_kfree);
        }
        // sQ / sdO may only be overwritten (next item) after the last MMAs finished
        mbar_wait(bar_kfree, (g + T - 1) & 1);
      }
      __syncwarp();
    } else {
      const int quarter = warp & 3, slice = warp >> 2;
      const int row = quarter * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      const int q = MQA ? q0 + (row >> 2) : q0 + row;
      const int h = MQA ? (row & 3) : hsel;
      if (HAS_BIAS && !SAVED) {
        const float4* src = P.geo + ((size_t)b * P.nQp + q0) * GEO_F4;
        for (int i = tid; i < QT * GEO_F4; i += NCOMPUTE) sGeo[i] = __ldg(src + i);
      }
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

      for (int j = 0; j < T; ++j) {
        const uint32_t gj = g + j;
        const int key0 = (tile_begin + j) * BN;
        if (HAS_BIAS && SAVED) mbar_wait(bar_k + (gj & 1), (gj >> 1) & 1);     // the bias tile has landed
        if (HAS_BIAS && !SAVED) {
          mbar_wait(bar_k + (gj & 1), (gj >> 1) & 1);
          const int kg = warp & 1;
          const float4 kx = sXyz[(gj & 1) * BN + kg * 32 + lane];
#pragma unroll 1
          for (int u = 0; u < 4; ++u) {
            const int qq = (warp >> 1) + 8 * u;
            sBias[qq * BIAS_STRIDE_F4 + kg * 32 + lane] =
                rpe_bias_pair(sGeo + qq * GEO_F4, kx.x, kx.y, kx.z, sTab, P.grid_n, P.log_scale, P.c1, P.c0);
          }
          named_bar_sync(1, NCOMPUTE);                     // (a) bias tile complete
        }
        mbar_wait(bar_s + (gj & 1), (gj >> 1) & 1);
        tc_fence_after();
        uint32_t sr[16], dr[16];
        tmem_ld16(tS0 + (gj & 1) * BN + lane_addr + slice * 16, sr);
        tmem_ld16(tdP0 + (gj & 1) * BN + lane_addr + slice * 16, dr);
        tmem_ld_wait();
        tc_fence_before();
        if (!SAVED) {
          if (HAS_BIAS) fence_proxy_async_smem();          // key xyz was read (generic proxy), bulk copies rewrite it (async proxy)
          mbar_arrive(bar_p + (gj & 1));                   // TMEM buffers of this tile may be overwritten
        }
        const float* brow = nullptr;
        if (HAS_BIAS)
          brow = reinterpret_cast<const float*>(sBias + (SAVED ? (gj & 1) * (QT * BIAS_STRIDE_F4) : 0) + (row >> 2) * BIAS_STRIDE_F4 +
                                                slice * 16) + (row & 3);
        const bool tail_tile = key0 + BN > P.nK;
        uint32_t pk[8], dk_[8];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float s = __uint_as_float(sr[c + e]);
            if (HAS_BIAS) s += brow[(c + e) * 4];
            float p = ex2_approx(s * LOG2E - lse2);
            if (tail_tile && key0 + slice * 16 + c + e >= P.nK) p = 0.f;       // only the last key tile is ragged
            const float ds = p * (__uint_as_float(dr[c + e]) - Drow);        // = g * dS
            pv[e] = p; dv[e] = ds;
          }
          pk[c >> 1] = pack_f16x2(pv[0], pv[1]);
          dk_[c >> 1] = pack_f16x2(dv[0], dv[1]);
        }
        if (SAVED) {
          // The bias tile was read through the generic proxy and will be overwritten through the async proxy (bulk
          // copies of tile gj + 2): the mbarrier arrive / wait chain alone does not order the two proxies.  Without this
          // fence dq / dk / dv differed run to run in 8-17 % of the launches of a 6-tile item (tests/dev_determinism.py;
          // 0 of 400 with it).
          fence_proxy_async_smem();
          mbar_arrive(bar_p + (gj & 1));                   // TMEM buffers and the bias buffer of this tile are free
        }
        {
          uint4* dstp = reinterpret_cast<uint4*>(P.pb + grow + key0 + slice * 16);
          uint4* dstd = reinterpret_cast<uint4*>(P.dsb + grow + key0 + slice * 16);
          dstp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]); dstp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          dstd[0] = make_uint4(dk_[0], dk_[1], dk_[2], dk_[3]); dstd[1] = make_uint4(dk_[4], dk_[5], dk_[6], dk_[7]);
        }
        if (HAS_BIAS && !SAVED) named_bar_sync(2, NCOMPUTE);   // (b) every thread has read its bias: the tile may be rewritten
      }
      named_bar_sync(2, NCOMPUTE);                         // sRow / sGeo reusable by the next item
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

// dq/dk/dv padded GEMM outputs -> user layouts
struct UnpackParams {
  int B, nQ, nK, nQp, nKp, kvh;
  const float *dqp, *dkp, *dvp;
  float *dq, *dk, *dv;
  const unsigned* absmax_bits;
};
__global__ void bwd_unpack_kernel(UnpackParams U) {
  const float ginv = 1.0f / vdetr_grad_scale(*U.absmax_bits);
  const size_t nq = (size_t)U.B * U.nQ * 4 * 64, nk = (size_t)U.B * U.nK * U.kvh * 64;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nq + 2 * nk; i += (size_t)gridDim.x * blockDim.x) {
    if (i < nq) {
      const int d = (int)(i & 63), h = (int)((i >> 6) & 3);
      const size_t bq = i >> 8;
      const int q = (int)(bq % U.nQ), b = (int)(bq / U.nQ);
      const size_t row = U.kvh == 1 ? ((size_t)b * U.nQp + q) * 4 + h : ((size_t)b * 4 + h) * U.nQp + q;
      U.dq[i] = U.dqp[row * 64 + d] * ginv;
    } else {
      const size_t e = (i - nq) % nk;
      const bool isv = (i - nq) >= nk;
      const int d = (int)(e & 63);
      const size_t r = e >> 6;
      const int hk = (int)(r % U.kvh);
      const int key = (int)((r / U.kvh) % U.nK), b = (int)(r / ((size_t)U.kvh * U.nK));
      const size_t src = (((size_t)b * U.kvh + hk) * U.nKp + key) * 64 + d;
      if (isv) U.dv[e] = U.dvp[src] * ginv; else U.dk[e] = U.dkp[src] * ginv;
    }
  }
}

struct BwdPlan {
  int nQp, nKp, mtiles, splits, tiles_per_split, items;
  size_t off_qp, off_dop, off_kp, off_vp, off_max, off_xyz, off_geo, off_pb, off_dsb, off_dt, off_dqp, off_dkp, off_dvp, total;
};
BwdPlan make_plan(const VdetrXattnShape* s) {
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
  p.off_dop = take(rows * 64 * 2);
  p.off_kp = take(krows * 64 * 2);
  p.off_vp = take(krows * 64 * 2);
  p.off_max = take(16);
  p.off_xyz = take(s->has_bias ? (size_t)s->B * p.nKp * 16 : 0);
  p.off_geo = take(s->has_bias ? (size_t)s->B * p.nQp * GEO_F4 * 16 : 0);
  p.off_pb = take(rows * p.nKp * 2);
  p.off_dsb = take(rows * p.nKp * 2);
  p.off_dt = take(s->has_bias ? rpe_dtables_scratch_bytes(s) : 0);
  p.off_dqp = take(rows * 64 * 4);
  p.off_dkp = take(krows * 64 * 4);
  p.off_dvp = take(krows * 64 * 4);
  p.total = o;
  return p;
}

cublasHandle_t get_cublas() {
  static cublasHandle_t h[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!h[dev] && cublasCreate(&h[dev]) != CUBLAS_STATUS_SUCCESS) return nullptr;
  return h[dev];
}

// row-major C[M,N] = op(A) op(B), fp16 inputs, fp32 output, strided batch
int gemm_rm(cublasHandle_t hnd, bool ta, bool tb, int M, int N, int K, const __half* A, int lda, long long sa,
            const __half* Bm, int ldb, long long sb, float* C, int ldc, long long sc, int batch) {
  const float alpha = 1.f, beta = 0.f;
  cublasStatus_t st = cublasGemmStridedBatchedEx(hnd, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K,
                                                 &alpha, Bm, CUDA_R_16F, ldb, sb, A, CUDA_R_16F, lda, sa, &beta, C, CUDA_R_32F,
                                                 ldc, sc, batch, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
  return st == CUBLAS_STATUS_SUCCESS ? 0 : VDETR_ERR_UNSUPPORTED;
}

}  // namespace

size_t tc_xattn_bwd_workspace(const VdetrXattnShape* s) { return make_plan(s).total; }

int tc_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                 const float* ref, const float* ang, const float* tables, const float* out, const float* lse,
                 const float* dout, const float* bias_saved, float* dq, float* dk, float* dv, float* dtables, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  const bool mqa = s->kv_heads == 1;
  if (s->has_bias && !mqa) return VDETR_ERR_UNSUPPORTED;
  const BwdPlan pl = make_plan(s);
  if (!ws || ws_bytes < pl.total || (reinterpret_cast<uintptr_t>(ws) & 255) != 0) return VDETR_ERR_WORKSPACE;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);

  VdetrPack pk = {};
  pk.B = s->B; pk.nQ = s->nQ; pk.nK = s->nK; pk.nQp = pl.nQp; pk.nKp = pl.nKp; pk.kvh = s->kv_heads; pk.has_bias = s->has_bias;
  pk.q = q; pk.k = k; pk.v = v; pk.xyz = xyz; pk.ref = ref; pk.ang = (s->has_bias && s->rotate) ? ang : nullptr; pk.dout = dout;
  unsigned* absmax = reinterpret_cast<unsigned*>(w + pl.off_max);
  VDETR_CUDA_TRY(cudaMemsetAsync(absmax, 0, 4, st));
  vdetr_absmax_kernel<<<vdetr_num_sms() * 2, 256, 0, st>>>(dout, (size_t)s->B * s->nQ * 4 * 64, absmax);
  VDETR_LAUNCH_CHECK();
  pk.dout_absmax_bits = absmax;
  pk.qp = reinterpret_cast<__half*>(w + pl.off_qp);
  pk.dop = reinterpret_cast<__half*>(w + pl.off_dop);
  pk.kp = reinterpret_cast<__half*>(w + pl.off_kp);
  pk.vp = reinterpret_cast<__half*>(w + pl.off_vp);
  pk.vtp = nullptr;
  pk.xyz4 = reinterpret_cast<float4*>(w + pl.off_xyz);
  pk.geo = reinterpret_cast<float4*>(w + pl.off_geo);
  vdetr_pack_kernel<<<vdetr_num_sms() * 4, 256, 0, st>>>(pk);
  VDETR_LAUNCH_CHECK();

  const uint64_t rows = (uint64_t)s->B * pl.nQp * 4, krows = (uint64_t)s->B * s->kv_heads * pl.nKp;
  CUtensorMap tmQ, tmdO, tmK, tmV;
  int rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmQ, pk.qp, rows, BM, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmdO, pk.dop, rows, BM, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmK, pk.kp, krows, BN, true))) return rc;
  if ((rc = vdetr_make_tmap_bf16_rows64(&tmV, pk.vp, krows, BN, true))) return rc;

  BwdParams P = {};
  P.B = s->B; P.nQ = s->nQ; P.nK = s->nK; P.nQp = pl.nQp; P.nKp = pl.nKp; P.kvh = s->kv_heads;
  P.mtiles = pl.mtiles; P.splits = pl.splits; P.tiles_per_split = pl.tiles_per_split; P.items = pl.items;
  P.grid_n = s->has_bias ? s->grid_n : 0;
  P.log_scale = s->log_scale;
  P.c1 = s->has_bias ? (float)s->grid_n / (2.0f * 3.0f * s->max_value) : 0.f;
  P.c0 = s->has_bias ? 0.5f * (float)(s->grid_n - 1) : 0.f;
  P.xyz4 = pk.xyz4; P.geo = pk.geo; P.tables = reinterpret_cast<const float4*>(tables);
  P.out = out; P.dout = dout; P.lse = lse;
  P.absmax_bits = absmax;
  P.pb = reinterpret_cast<__half*>(w + pl.off_pb);
  P.dsb = reinterpret_cast<__half*>(w + pl.off_dsb);
  const bool saved = s->has_bias && bias_saved != nullptr;
  P.bias_in = saved ? reinterpret_cast<const float4*>(bias_saved) : nullptr;

  const int table_bytes = (s->has_bias && !saved) ? rpe::pair_table_bytes(s->grid_n) : 0;
  const SmemLayout L = smem_layout(table_bytes, saved ? 2 : 1);
  if (L.total + 1024 > 232448) return VDETR_ERR_UNSUPPORTED;
  const size_t smem = L.total + 1024;
  const int grid = pl.items < vdetr_num_sms() ? pl.items : vdetr_num_sms();
  {
  VdetrTimingScope timing(s->has_bias ? VDETR_T_BWD : VDETR_T_COUNT, st);
  if (saved) {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_bwd_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpe_xattn_bwd_kernel<true, true, true><<<grid, NTHREADS, smem, st>>>(tmQ, tmdO, tmK, tmV, P);
  } else if (s->has_bias) {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_bwd_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpe_xattn_bwd_kernel<true, true, false><<<grid, NTHREADS, smem, st>>>(tmQ, tmdO, tmK, tmV, P);
  } else if (mqa) {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_bwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpe_xattn_bwd_kernel<false, true, false><<<grid, NTHREADS, smem, st>>>(tmQ, tmdO, tmK, tmV, P);
  } else {
    VDETR_CUDA_TRY(cudaFuncSetAttribute(rpe_xattn_bwd_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rpe_xattn_bwd_kernel<false, false, false><<<grid, NTHREADS, smem, st>>>(tmQ, tmdO, tmK, tmV, P);
  }
  }
  VDETR_LAUNCH_CHECK();

  // ---- pass 2: dQ = dS K, dK = dS^T Q, dV = P^T dO
  cublasHandle_t hnd = get_cublas();
  if (!hnd) return VDETR_ERR_UNSUPPORTED;
  if (cublasSetStream(hnd, st) != CUBLAS_STATUS_SUCCESS) return VDETR_ERR_UNSUPPORTED;
  float* dqp = reinterpret_cast<float*>(w + pl.off_dqp);
  float* dkp = reinterpret_cast<float*>(w + pl.off_dkp);
  float* dvp = reinterpret_cast<float*>(w + pl.off_dvp);
  const int batch = mqa ? s->B : s->B * 4;
  const int rpb = mqa ? pl.nQp * 4 : pl.nQp;              // attention rows per batch entry
  const long long sRows = (long long)rpb * pl.nKp, sRow64 = (long long)rpb * 64, sK64 = (long long)pl.nKp * 64;
  if ((rc = gemm_rm(hnd, false, false, rpb, 64, pl.nKp, P.dsb, pl.nKp, sRows, pk.kp, 64, sK64, dqp, 64, sRow64, batch))) return rc;
  if ((rc = gemm_rm(hnd, true, false, pl.nKp, 64, rpb, P.dsb, pl.nKp, sRows, pk.qp, 64, sRow64, dkp, 64, sK64, batch))) return rc;
  if ((rc = gemm_rm(hnd, true, false, pl.nKp, 64, rpb, P.pb, pl.nKp, sRows, pk.dop, 64, sRow64, dvp, 64, sK64, batch))) return rc;
  UnpackParams U = {s->B, s->nQ, s->nK, pl.nQp, pl.nKp, s->kv_heads, dqp, dkp, dvp, dq, dk, dv, absmax};
  bwd_unpack_kernel<<<vdetr_num_sms() * 4, 256, 0, st>>>(U);
  VDETR_LAUNCH_CHECK();

  // ---- pass 3: dTables
  if (s->has_bias) {
    if ((rc = rpe_dtables_launch(s, pl.nQp, pl.nKp, pk.xyz4, pk.geo, P.dsb, absmax, 0, dtables, w + pl.off_dt,
                                 rpe_dtables_scratch_bytes(s), st)))
      return rc;
  }
  return 0;
}
