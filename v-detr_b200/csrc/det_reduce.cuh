// Deterministic cross-CTA column sums for the BatchNorm / LayerNorm / bias-gradient kernels.
//
// Float atomics make a sum depend on the order in which CTAs happen to arrive: the BatchNorm statistics of two identical
// forward passes differed in the last bit, which is enough to flip the decoder's top-k proposal selection (stock
// nn.BatchNorm1d, which the reference relies on -- models/helpers.py:24-28 -- is deterministic).  Here every CTA stores
// its partial column sums, takes a ticket, and the CTA that draws the last ticket adds the partials IN BLOCK ORDER:
// the result does not depend on which CTA is last.
#pragma once
#include <cuda_runtime.h>

constexpr int VDETR_RED_MAX_BLOCKS = 128;      // partial sums per reduction (grids of the reducing kernels are capped to this)

// workspace of one reduction over `cols` columns of two quantities: [2 * cols] results, [128][2 * cols] partials, 32 tickets,
// then (SyncBatchNorm only) [32] scalars and [2 * cols] sums over all ranks
static inline size_t vdetr_reduce_ws_floats(int cols) { return (size_t)(VDETR_RED_MAX_BLOCKS + 1) * 2 * cols + 32 + 32 + 2 * (size_t)cols; }

// Called by ALL threads of every CTA after the CTA has written part[blockIdx.x * ncols + c] for all c.
// ncols columns; results go to out_a[c] (c < split) or out_b[c - split].  *ticket must be 0 before the launch.
template <int NT>
__device__ __forceinline__ void det_finish_columns(const float* part, int nblocks, int ncols, unsigned* ticket, float* out_a,
                                                   float* out_b, int split) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)(nblocks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < ncols; c += NT) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    int b = 0;
    for (; b + 7 < nblocks; b += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += __ldcg(part + (size_t)(b + j) * ncols + c);
    }
    for (int j = 0; b < nblocks; ++b, ++j) s[j] += __ldcg(part + (size_t)b * ncols + c);
    const float t = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    if (c < split) out_a[c] = t; else out_b[c - split] = t;
  }
}
