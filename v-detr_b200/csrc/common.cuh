// Shared helpers for libvdetr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/vdetr_b200.h"

#define VDETR_CUDA_TRY(expr)                          \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return (int)_e;            \
  } while (0)

// every kernel launch of the library is followed by this check; it also counts the launches (vdetr_launch_count)
extern unsigned long long g_vdetr_launches;
#define VDETR_LAUNCH_CHECK()                          \
  do {                                                \
    ++g_vdetr_launches;                               \
    cudaError_t _e = cudaGetLastError();              \
    if (_e != cudaSuccess) return (int)_e;            \
  } while (0)

static inline size_t vdetr_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static inline int vdetr_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------------------
// Optional kernel timing (bench.py): CUDA events recorded on the launch stream around the dominant
// kernels.  Disabled by default; enabling costs two cudaEventRecord per launch.
enum VdetrTimedKernel { VDETR_T_FWD = 0, VDETR_T_BWD = 1, VDETR_T_DTABLES = 2, VDETR_T_BWD2 = 3, VDETR_T_COUNT = 4 };
struct VdetrTimingScope {
  int kind;
  cudaStream_t st;
  cudaEvent_t stop;
  bool on;
  VdetrTimingScope(int kind, cudaStream_t st);
  ~VdetrTimingScope();
};
