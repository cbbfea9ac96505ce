#include "common.cuh"
#include "rpe_internal.h"
size_t tc_xattn_bwd_workspace(const VdetrXattnShape*) { return 0; }
int tc_xattn_bwd(const VdetrXattnShape*, const float*, const float*, const float*, const float*, const float*, const float*,
                 const float*, const float*, const float*, const float*, float*, float*, float*, float*, void*, size_t,
                 cudaStream_t) { return VDETR_ERR_UNSUPPORTED; }
