// Batched rectangular linear sum assignment on the device: the Hungarian matching of DETR-style set prediction
// (/root/reference/criterion.py:205-228 calls scipy.optimize.linear_sum_assignment per scene on a [nQ x nactual_gt] cost
// matrix after a .cpu() round trip, nine times per training step).  One CTA per scene runs the same algorithm scipy does
// -- shortest augmenting paths with dual variables (Jonker-Volgenant / Crouse 2016), on the transposed problem: rows = ground
// truth boxes, columns = queries, nr <= nc -- with the per-column state in registers and a block-wide arg-min per
// path step.  Arithmetic on the duals is FP64 like scipy's.  Exact ties between candidate columns may be broken differently
// from scipy (its order depends on the history of an index array); the total cost is the same.
#include <float.h>
#include "common.cuh"

namespace {

constexpr int LS_THREADS = 1024;
constexpr int LS_CPT = 4;                 // columns (queries) per thread: nQ <= 4096
constexpr int LS_MAX_ROWS = 512;          // ground-truth boxes per scene

struct Cand {
  double val;
  int key;      // (row4col[j] >= 0) << 30 | j : smaller wins on equal val (unassigned columns first, as scipy prefers)
};
__device__ __forceinline__ Cand cand_min(Cand a, Cand b) { return (b.val < a.val || (b.val == a.val && b.key < a.key)) ? b : a; }

__global__ void __launch_bounds__(LS_THREADS, 1)
lsap_kernel(const float* __restrict__ cost, const int32_t* __restrict__ nactual, int nQ, int ngt, long long* __restrict__ inds,
            float* __restrict__ mask) {
  extern __shared__ int s_int[];
  int* row4col = s_int;                 // [nQ]
  int* path = s_int + nQ;               // [nQ]
  __shared__ double u[LS_MAX_ROWS];
  __shared__ int col4row[LS_MAX_ROWS];
  __shared__ Cand red[32];
  __shared__ Cand s_best;
  __shared__ int s_next;                // row to scan next, or -1 - sink

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* C = cost + (size_t)b * nQ * ngt;
  int nr = nactual ? __ldg(nactual + b) : ngt;
  nr = max(0, min(nr, min(ngt, nQ)));
  double v[LS_CPT], sp[LS_CPT];
  bool sc[LS_CPT];
#pragma unroll
  for (int k = 0; k < LS_CPT; ++k) v[k] = 0.0;
  for (int j = tid; j < nQ; j += LS_THREADS) row4col[j] = -1;
  for (int i = tid; i < nr; i += LS_THREADS) { u[i] = 0.0; col4row[i] = -1; }
  __syncthreads();

  for (int cur = 0; cur < nr; ++cur) {
#pragma unroll
    for (int k = 0; k < LS_CPT; ++k) { sp[k] = DBL_MAX; sc[k] = false; }
    double minVal = 0.0;
    int i = cur, sink = -1;
    while (sink < 0) {
      const double ui = u[i];
      Cand best = {DBL_MAX, 0x7FFFFFFF};
#pragma unroll
      for (int k = 0; k < LS_CPT; ++k) {
        const int j = tid + k * LS_THREADS;
        if (j < nQ && !sc[k]) {
          const double r = minVal + (double)__ldg(C + (size_t)j * ngt + i) - ui - v[k];
          if (r < sp[k]) { sp[k] = r; path[j] = i; }
          best = cand_min(best, Cand{sp[k], ((row4col[j] >= 0) << 30) | j});
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        Cand other;
        other.val = __shfl_xor_sync(0xffffffffu, best.val, o);
        other.key = __shfl_xor_sync(0xffffffffu, best.key, o);
        best = cand_min(best, other);
      }
      if (lane == 0) red[warp] = best;
      __syncthreads();
      if (warp == 0) {
        Cand c = red[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          Cand other;
          other.val = __shfl_xor_sync(0xffffffffu, c.val, o);
          other.key = __shfl_xor_sync(0xffffffffu, c.key, o);
          c = cand_min(c, other);
        }
        if (lane == 0) {
          s_best = c;
          const int jm = c.key & 0x3FFFFFFF;
          s_next = (c.val == DBL_MAX) ? -1 - nQ : (row4col[jm] < 0 ? -1 - jm : row4col[jm]);
        }
      }
      __syncthreads();
      const Cand win = s_best;
      const int nxt = s_next;
      if (nxt == -1 - nQ) { sink = nQ; break; }            // infeasible (non-finite costs): leave the row unmatched
      minVal = win.val;
      const int jm = win.key & 0x3FFFFFFF;
      if ((jm % LS_THREADS) == tid) sc[jm / LS_THREADS] = true;
      if (nxt < 0) sink = -1 - nxt; else i = nxt;
      // (the next iteration's barrier protects red / s_best / s_next)
    }
    if (sink < nQ) {
      // dual update: u[cur] += minVal; rows reached through a scanned column j: u[row4col[j]] += minVal - sp[j]; v[j] -= ...
      if (tid == 0) u[cur] += minVal;
#pragma unroll
      for (int k = 0; k < LS_CPT; ++k) {
        const int j = tid + k * LS_THREADS;
        if (j < nQ && sc[k]) {
          const double d = minVal - sp[k];
          v[k] -= d;
          const int r = row4col[j];
          if (r >= 0) u[r] += d;
        }
      }
      __syncthreads();
      if (tid == 0) {                                       // augment along the alternating path
        int j = sink, r;
        do {
          r = path[j];
          row4col[j] = r;
          const int t = col4row[r];
          col4row[r] = j;
          j = t;
        } while (r != cur);
      }
    }
    __syncthreads();
  }
  for (int j = tid; j < nQ; j += LS_THREADS) {
    const int r = row4col[j];
    inds[(size_t)b * nQ + j] = r >= 0 ? r : 0;
    mask[(size_t)b * nQ + j] = r >= 0 ? 1.0f : 0.0f;
  }
}

}  // namespace

extern "C" int vdetr_lsap(const float* cost, const int32_t* nactual_gt, int B, int nQ, int ngt, long long* per_prop_gt_inds,
                          float* proposal_matched_mask, void* stream) {
  if (B < 0 || nQ < 0 || ngt < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || nQ == 0) return 0;
  if (!per_prop_gt_inds || !proposal_matched_mask || (ngt > 0 && !cost)) return VDETR_ERR_BAD_ARG;
  if (nQ > LS_THREADS * LS_CPT || ngt > LS_MAX_ROWS || ngt > nQ) return VDETR_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (ngt == 0) {
    VDETR_CUDA_TRY(cudaMemsetAsync(per_prop_gt_inds, 0, (size_t)B * nQ * sizeof(long long), st));
    VDETR_CUDA_TRY(cudaMemsetAsync(proposal_matched_mask, 0, (size_t)B * nQ * sizeof(float), st));
    return 0;
  }
  const size_t smem = (size_t)2 * nQ * sizeof(int);
  lsap_kernel<<<B, LS_THREADS, smem, st>>>(cost, nactual_gt, nQ, ngt, per_prop_gt_inds, proposal_matched_mask);
  VDETR_LAUNCH_CHECK();
  return 0;
}
