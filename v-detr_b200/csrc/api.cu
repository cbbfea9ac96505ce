// C ABI dispatch for libvdetr_b200 (see include/vdetr_b200.h).
#include "common.cuh"
#include "rpe_internal.h"

namespace {
constexpr int TIMING_RING = 512;
struct TimingState {
  bool enabled = false;
  cudaEvent_t start[VDETR_T_COUNT][TIMING_RING];
  cudaEvent_t stop[VDETR_T_COUNT][TIMING_RING];
  int n[VDETR_T_COUNT] = {0, 0, 0, 0};
  bool created = false;
} g_timing;
}  // namespace

VdetrTimingScope::VdetrTimingScope(int k, cudaStream_t s) : kind(k), st(s), stop(nullptr), on(false) {
  if (k < 0 || k >= VDETR_T_COUNT || !g_timing.enabled || g_timing.n[k] >= TIMING_RING) return;
  const int i = g_timing.n[k]++;
  cudaEventRecord(g_timing.start[k][i], st);
  stop = g_timing.stop[k][i];
  on = true;
}
VdetrTimingScope::~VdetrTimingScope() {
  if (on) cudaEventRecord(stop, st);
}

unsigned long long g_vdetr_launches = 0;

extern "C" {

unsigned long long vdetr_launch_count(int reset) {
  const unsigned long long n = g_vdetr_launches;
  if (reset) g_vdetr_launches = 0;
  return n;
}

int vdetr_timing_enable(int enable) {
  if (enable && !g_timing.created) {
    for (int k = 0; k < VDETR_T_COUNT; ++k)
      for (int i = 0; i < TIMING_RING; ++i) {
        VDETR_CUDA_TRY(cudaEventCreate(&g_timing.start[k][i]));
        VDETR_CUDA_TRY(cudaEventCreate(&g_timing.stop[k][i]));
      }
    g_timing.created = true;
  }
  g_timing.enabled = enable != 0;
  for (int k = 0; k < VDETR_T_COUNT; ++k) g_timing.n[k] = 0;
  return 0;
}

int vdetr_timing_read(float* total_ms /*[4]*/, int* launches /*[4]*/) {
  for (int k = 0; k < VDETR_T_COUNT; ++k) {
    float sum = 0.f;
    for (int i = 0; i < g_timing.n[k]; ++i) {
      VDETR_CUDA_TRY(cudaEventSynchronize(g_timing.stop[k][i]));
      float ms = 0.f;
      VDETR_CUDA_TRY(cudaEventElapsedTime(&ms, g_timing.start[k][i], g_timing.stop[k][i]));
      sum += ms;
    }
    total_ms[k] = sum;
    launches[k] = g_timing.n[k];
    g_timing.n[k] = 0;
  }
  return 0;
}

const char* vdetr_version(void) { return "vdetr_b200 0.1 sm_100a"; }

const char* vdetr_error_string(int code) {
  switch (code) {
    case VDETR_OK: return "ok";
    case VDETR_ERR_BAD_ARG: return "vdetr_b200: bad argument";
    case VDETR_ERR_UNSUPPORTED: return "vdetr_b200: unsupported shape or option";
    case VDETR_ERR_WORKSPACE: return "vdetr_b200: workspace missing or too small";
    case VDETR_ERR_NO_DRIVER: return "vdetr_b200: CUDA driver entry point unavailable";
    default: return cudaGetErrorString((cudaError_t)code);
  }
}

size_t vdetr_xattn_fwd_workspace_bytes(const VdetrXattnShape* s, int impl) {
  if (vdetr_check_shape(s) != 0) return 0;
  return impl == 0 ? tc_xattn_fwd_workspace(s) : 0;
}

size_t vdetr_xattn_bias_save_bytes(const VdetrXattnShape* s, int impl) {
  if (vdetr_check_shape(s) != 0 || impl != 0) return 0;
  return tc_xattn_bias_save_bytes(s);
}

int vdetr_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                    const float* ref_pts, const float* ref_angle, const float* tables, float* out, float* lse,
                    float* bias_save, float dropout_p, const uint64_t* dropout_seed, void* workspace, size_t workspace_bytes,
                    int impl, void* stream) {
  int rc = vdetr_check_shape(s);
  if (rc) return rc;
  if (s->B == 0 || s->nQ == 0) return 0;
  if (s->nK == 0) return VDETR_ERR_BAD_ARG;
  if (!q || !k || !v || !out || !lse) return VDETR_ERR_BAD_ARG;
  if (s->has_bias && (!xyz || !ref_pts || !tables)) return VDETR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 1) {
    if (dropout_p > 0.f) return VDETR_ERR_UNSUPPORTED;      // the validation kernels implement no dropout
    return simt_xattn_fwd(s, q, k, v, xyz, ref_pts, ref_angle, tables, out, lse, st);
  }
  if (impl == 0)
    return tc_xattn_fwd(s, q, k, v, xyz, ref_pts, ref_angle, tables, out, lse, bias_save, dropout_p,
                        reinterpret_cast<const unsigned long long*>(dropout_seed), workspace, workspace_bytes, st);
  return VDETR_ERR_BAD_ARG;
}

size_t vdetr_xattn_bwd_workspace_bytes(const VdetrXattnShape* s, int impl, int bias_is_saved) {
  if (vdetr_check_shape(s) != 0) return 0;
  return impl == 0 ? tc_xattn_bwd_workspace(s, bias_is_saved) : 0;
}

int vdetr_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                    const float* ref_pts, const float* ref_angle, const float* tables, const float* out,
                    const float* lse, const float* dout, const float* bias_saved, float dropout_p,
                    const uint64_t* dropout_seed, float* dq, float* dk, float* dv, float* dtables, void* workspace,
                    size_t workspace_bytes, int impl, void* stream) {
  int rc = vdetr_check_shape(s);
  if (rc) return rc;
  if (s->B == 0 || s->nQ == 0 || s->nK == 0) return (s->nK == 0 && s->B && s->nQ) ? VDETR_ERR_BAD_ARG : 0;
  if (!q || !k || !v || !out || !lse || !dout || !dq || !dk || !dv) return VDETR_ERR_BAD_ARG;
  if (s->has_bias && (!xyz || !ref_pts || !tables || !dtables)) return VDETR_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 1) {
    if (dropout_p > 0.f) return VDETR_ERR_UNSUPPORTED;
    return simt_xattn_bwd(s, q, k, v, xyz, ref_pts, ref_angle, tables, out, lse, dout, dq, dk, dv, dtables, st);
  }
  if (impl == 0)
    return tc_xattn_bwd(s, q, k, v, xyz, ref_pts, ref_angle, tables, out, lse, dout, bias_saved, dropout_p,
                        reinterpret_cast<const unsigned long long*>(dropout_seed), dq, dk, dv, dtables, workspace,
                        workspace_bytes, st);
  return VDETR_ERR_BAD_ARG;
}

int vdetr_rpe_bias(const VdetrXattnShape* s, const float* xyz, const float* ref_pts, const float* ref_angle,
                   const float* tables, float* rpe, void* stream) {
  int rc = vdetr_check_shape(s);
  if (rc) return rc;
  if (!s->has_bias || !xyz || !ref_pts || !tables || !rpe) return VDETR_ERR_BAD_ARG;
  return rpe_bias_launch(s, xyz, ref_pts, ref_angle, tables, rpe, (cudaStream_t)stream);
}

size_t vdetr_rpe_dtables_workspace_bytes(const VdetrXattnShape* s) {
  if (vdetr_check_shape(s) != 0 || !s->has_bias) return 0;
  return rpe_dtables_workspace(s);
}

int vdetr_rpe_dtables(const VdetrXattnShape* s, const float* xyz, const float* ref_pts, const float* ref_angle,
                      const float* dbias, float* dtables, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = vdetr_check_shape(s);
  if (rc) return rc;
  if (!s->has_bias || !xyz || !ref_pts || !dbias || !dtables) return VDETR_ERR_BAD_ARG;
  return rpe_dtables_dense(s, xyz, ref_pts, ref_angle, dbias, dtables, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
