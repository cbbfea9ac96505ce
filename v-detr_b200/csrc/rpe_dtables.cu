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
