// Box decode of one decoder level in one kernel per direction (models/vdetr_transformer.py:244-333 with
// num_angle_bin = 1, i.e. angle == 0: the ScanNet configuration; other configurations keep the PyTorch ops).
//
//   pre_center = pre_center_normalized * scene + lo            pre_size = pre_size_normalized * scene
//   center     = center_reg * pre_size + pre_center            center_normalized = (center - lo) / scene
//   size       = exp(size_reg) * pre_size                      size_normalized   = size / scene
//   corners    = dataset_config.box_parametrization_to_corners(center, size, 0)   (camera frame, utils/box_util.py:294-358)
//   ref_lidar  = convert_corners_camera2lidar(corners)         (the next layer's reference_point, :98-102, no gradient)
// The reference spends ~35 elementwise launches per level on this (x 9 levels, forward and backward).
#include "common.cuh"

namespace {

struct BoxDecodeParams {
  int B, nQ;
  const float *center_reg, *size_reg, *pre_cn, *pre_sn, *lo, *hi;        // [B,nQ,3] x4, [B,3] x2
  float *center, *center_norm, *size, *size_norm, *pre_center, *pre_size;  // [B,nQ,3]
  float *corners, *ref_lidar;                                              // [B,nQ,8,3]
};

__device__ __constant__ float kSx[8] = {1, 1, -1, -1, 1, 1, -1, -1};
__device__ __constant__ float kSy[8] = {1, 1, 1, 1, -1, -1, -1, -1};
__device__ __constant__ float kSz[8] = {1, -1, -1, 1, 1, -1, -1, 1};

__global__ void box_decode_fwd_kernel(const BoxDecodeParams P) {
  const int T = P.B * P.nQ;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int b = t / P.nQ;
    float c[3], s[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float lo = __ldg(P.lo + b * 3 + a), scene = __ldg(P.hi + b * 3 + a) - lo;
      const float pc = P.pre_cn[t * 3 + a] * scene + lo, ps = P.pre_sn[t * 3 + a] * scene;
      c[a] = P.center_reg[t * 3 + a] * ps + pc;
      s[a] = expf(P.size_reg[t * 3 + a]) * ps;
      P.pre_center[t * 3 + a] = pc; P.pre_size[t * 3 + a] = ps;
      P.center[t * 3 + a] = c[a]; P.center_norm[t * 3 + a] = (c[a] - lo) / scene;
      P.size[t * 3 + a] = s[a]; P.size_norm[t * 3 + a] = s[a] / scene;
    }
    // camera frame: centre (x, -z, y); half extents l = size.x (x), h = size.z (y), w = size.y (z)
    const float cx = c[0], cy = -c[2], cz = c[1];
    const float hl = s[0] * 0.5f, hw = s[1] * 0.5f, hh = s[2] * 0.5f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = hl * kSx[i] + cx, y = hh * kSy[i] + cy, z = hw * kSz[i] + cz;
      float* co = P.corners + ((size_t)t * 8 + i) * 3;
      co[0] = x; co[1] = y; co[2] = z;
      float* rl = P.ref_lidar + ((size_t)t * 8 + i) * 3;       // (x_c, z_c, -y_c)
      rl[0] = x; rl[1] = z; rl[2] = -y;
    }
  }
}

struct BoxDecodeGradParams {
  int B, nQ;
  const float *size, *pre_size, *lo, *hi;                                   // saved by the forward
  const float *g_center, *g_center_norm, *g_size, *g_size_norm, *g_corners;  // any may be null (no gradient)
  float *g_center_reg, *g_size_reg;                                         // [B,nQ,3]
};

__global__ void box_decode_bwd_kernel(const BoxDecodeGradParams P) {
  const int T = P.B * P.nQ;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int b = t / P.nQ;
    float gc[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f};
    if (P.g_corners) {
      float sx = 0.f, sy = 0.f, sz = 0.f, wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* g = P.g_corners + ((size_t)t * 8 + i) * 3;
        const float gx = g[0], gy = g[1], gz = g[2];
        sx += gx; sy += gy; sz += gz;
        wx += kSx[i] * gx; wy += kSy[i] * gy; wz += kSz[i] * gz;
      }
      gc[0] = sx; gc[2] = -sy; gc[1] = sz;                     // camera (x, y, z) = (c.x, -c.z, c.y)
      gs[0] = 0.5f * wx; gs[2] = 0.5f * wy; gs[1] = 0.5f * wz; // half extents (l, h, w) = (size.x, size.z, size.y) / 2
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float scene = __ldg(P.hi + b * 3 + a) - __ldg(P.lo + b * 3 + a);
      float dc = gc[a], ds = gs[a];
      if (P.g_center) dc += P.g_center[t * 3 + a];
      if (P.g_center_norm) dc += P.g_center_norm[t * 3 + a] / scene;
      if (P.g_size) ds += P.g_size[t * 3 + a];
      if (P.g_size_norm) ds += P.g_size_norm[t * 3 + a] / scene;
      P.g_center_reg[t * 3 + a] = dc * P.pre_size[t * 3 + a];
      P.g_size_reg[t * 3 + a] = ds * P.size[t * 3 + a];
    }
  }
}

}  // namespace

extern "C" {

int vdetr_box_decode_fwd(const float* center_reg, const float* size_reg, const float* pre_center_normalized,
                         const float* pre_size_normalized, const float* dims_min, const float* dims_max, int B, int nQ,
                         float* center, float* center_normalized, float* size, float* size_normalized, float* pre_center,
                         float* pre_size, float* corners, float* ref_lidar, void* stream) {
  if (B < 0 || nQ < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || nQ == 0) return 0;
  if (!center_reg || !size_reg || !pre_center_normalized || !pre_size_normalized || !dims_min || !dims_max || !center ||
      !center_normalized || !size || !size_normalized || !pre_center || !pre_size || !corners || !ref_lidar)
    return VDETR_ERR_BAD_ARG;
  BoxDecodeParams P = {B, nQ, center_reg, size_reg, pre_center_normalized, pre_size_normalized, dims_min, dims_max,
                       center, center_normalized, size, size_normalized, pre_center, pre_size, corners, ref_lidar};
  const int T = B * nQ;
  box_decode_fwd_kernel<<<(T + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P);
  VDETR_LAUNCH_CHECK();
  return 0;
}

int vdetr_box_decode_bwd(const float* size, const float* pre_size, const float* dims_min, const float* dims_max,
                         const float* g_center, const float* g_center_normalized, const float* g_size,
                         const float* g_size_normalized, const float* g_corners, int B, int nQ, float* g_center_reg,
                         float* g_size_reg, void* stream) {
  if (B < 0 || nQ < 0) return VDETR_ERR_BAD_ARG;
  if (B == 0 || nQ == 0) return 0;
  if (!size || !pre_size || !dims_min || !dims_max || !g_center_reg || !g_size_reg) return VDETR_ERR_BAD_ARG;
  BoxDecodeGradParams P = {B, nQ, size, pre_size, dims_min, dims_max, g_center, g_center_normalized, g_size, g_size_normalized,
                           g_corners, g_center_reg, g_size_reg};
  const int T = B * nQ;
  box_decode_bwd_kernel<<<(T + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P);
  VDETR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
