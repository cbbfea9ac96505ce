// Internal (non-ABI) entry points shared between the translation units of libvdetr_b200.
#pragma once
#include <cuda_runtime.h>
#include "../../include/vdetr_b200.h"

int vdetr_check_shape(const VdetrXattnShape* s);

// impl = 1 (validation kernels, rpe_simt.cu)
int simt_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                   const float* ref, const float* ang, const float* tables, float* out, float* lse, cudaStream_t st);
int simt_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                   const float* ref, const float* ang, const float* tables, const float* out, const float* lse,
                   const float* dout, float* dq, float* dk, float* dv, float* dtables, cudaStream_t st);
int rpe_bias_launch(const VdetrXattnShape* s, const float* xyz, const float* ref, const float* ang, const float* tables,
                    float* rpe, cudaStream_t st);

// impl = 0 (product kernels: tcgen05 + TMA, rpe_xattn_fwd.cu / rpe_xattn_bwd.cu)
size_t tc_xattn_fwd_workspace(const VdetrXattnShape* s);
int tc_xattn_fwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                 const float* ref, const float* ang, const float* tables, float* out, float* lse, void* ws, size_t ws_bytes,
                 cudaStream_t st);
size_t tc_xattn_bwd_workspace(const VdetrXattnShape* s);
int tc_xattn_bwd(const VdetrXattnShape* s, const float* q, const float* k, const float* v, const float* xyz,
                 const float* ref, const float* ang, const float* tables, const float* out, const float* lse,
                 const float* dout, float* dq, float* dk, float* dv, float* dtables, void* ws, size_t ws_bytes,
                 cudaStream_t st);
