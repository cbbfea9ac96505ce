/* CPU oracle: literal C restatement of the reference pointnet2 CUDA kernels (TEST INFRASTRUCTURE ONLY --
 * nothing under v-detr_b200/ may link or call this; it is the checker for tests/, smoke() and the
 * cpu_baseline leg of bench.py).
 *
 * Parity status: the reference ships no golden vectors for these ops and its extension has no CPU path
 * (AT_ASSERT(false, "CPU not supported"), src/sampling.cpp:84).  This file is pinned against the
 * reference extension itself, rebuilt for sm_100a as oracle/_ref/pn2_ref_ext.so (oracle/Makefile) and run on
 * the GPU box: tests/test_pointnet2_gpu.py::test_reference_ext_agrees_with_c_oracle, and against fixtures
 * captured from that run (tests/golden/pn2_ref_*.npz) on CPU.
 *
 * The emulation is deliberately literal (one loop per reference thread, the same shared-memory tree) so
 * that its tie-breaking is the reference's by construction rather than by argument.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (fmaf() is written out where nvcc contracts; see below).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* include/cuda_utils.h:15-21 */
static int opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

/* nvcc (-fmad=true) contracts  x*x + y*y + z*z  into  fma(z,z, fma(x,x, y*y)):  the SASS of the rebuilt
 * reference (oracle/_ref, cuobjdump) is  FMUL y,y ; FFMA x,x ; FFMA z,z  in the FPS kernel (both for |p|^2 and
 * for the distance) and in query_ball_point_kernel.  Verified on a B200: with any other order the indices
 * diverge from the reference after a few dozen rounds on lattice data. */
static inline float sumsq3(float x, float y, float z) { return fmaf(z, z, fmaf(x, x, y * y)); }

/* src/sampling_gpu.cu:72-176 + host init src/sampling.cpp:67-88 */
void pn2_ref_fps(int b, int n, int m, const float* dataset, int32_t* idxs) {
  if (m <= 0) return;
  const int S = opt_n_threads(n);
  float* temp = (float*)malloc(sizeof(float) * (size_t)n);
  float* dists = (float*)malloc(sizeof(float) * (size_t)S);
  int* dists_i = (int*)malloc(sizeof(int) * (size_t)S);
  for (int bi = 0; bi < b; ++bi) {
    const float* pts = dataset + (size_t)bi * n * 3;
    int32_t* out = idxs + (size_t)bi * m;
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    for (int j = 0; j < m; ++j) out[j] = 0;            /* torch::zeros */
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
      for (int tid = 0; tid < S; ++tid) {
        int besti = 0;
        float best = -1;
        for (int k = tid; k < n; k += S) {
          const float x2 = pts[k * 3 + 0], y2 = pts[k * 3 + 1], z2 = pts[k * 3 + 2];
          const float mag = sumsq3(x2, y2, z2);
          if (mag <= 1e-3) continue;                      /* float promoted to double vs 1e-3 */
          const float d = sumsq3(x2 - x1, y2 - y1, z2 - z1);
          const float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int stride = S / 2; stride >= 1; stride /= 2) {   /* the unrolled tree at :118-171 */
        for (int tid = 0; tid < stride; ++tid) {
          const float v1 = dists[tid], v2 = dists[tid + stride];
          const int i1 = dists_i[tid], i2 = dists_i[tid + stride];
          dists[tid] = v1 > v2 ? v1 : v2;                     /* max(v1, v2) */
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
  }
  free(temp); free(dists); free(dists_i);
}

/* src/ball_query_gpu.cu:12-47 + src/ball_query.cpp:22-24 (zeros) */
void pn2_ref_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                        int32_t* idx) {
  const float radius2 = radius * radius;
  memset(idx, 0, sizeof(int32_t) * (size_t)b * m * nsample);
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < m; ++j) {
      const float* c = new_xyz + ((size_t)bi * m + j) * 3;
      int32_t* row = idx + ((size_t)bi * m + j) * nsample;
      for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
        const float* p = xyz + ((size_t)bi * n + k) * 3;
        const float d2 = sumsq3(c[0] - p[0], c[1] - p[1], c[2] - p[2]);
        if (d2 < radius2) {
          if (cnt == 0) for (int l = 0; l < nsample; ++l) row[l] = k;
          row[cnt] = k;
          ++cnt;
        }
      }
    }
}

/* src/sampling_gpu.cu:11-23 / :37-50 */
void pn2_ref_gather(int b, int c, int n, int m, const float* points, const int32_t* idx, float* out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
}
void pn2_ref_gather_grad(int b, int c, int n, int m, const float* grad_out, const int32_t* idx, float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        grad_points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] += grad_out[((size_t)i * c + l) * m + j];
}

/* src/group_points_gpu.cu:11-31 / :46-67 */
void pn2_ref_group(int b, int c, int n, int npoints, int nsample, const float* points, const int32_t* idx, float* out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k)
          out[(((size_t)i * c + l) * npoints + j) * nsample + k] =
              points[((size_t)i * c + l) * n + idx[((size_t)i * npoints + j) * nsample + k]];
}
void pn2_ref_group_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int32_t* idx,
                        float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k)
          grad_points[((size_t)i * c + l) * n + idx[((size_t)i * npoints + j) * nsample + k]] +=
              grad_out[(((size_t)i * c + l) * npoints + j) * nsample + k];
}
