// Counter-based RNG for the attention dropout of the fused kernels (nn.Dropout on the attention probabilities,
// /root/reference/models/vdetr_transformer.py:751-752).  Philox4x32-10 (Salmon et al., SC'11): the mask of an
// element depends only on (seed, packed attention row, key index), so the forward and the backward kernels
// regenerate identical masks without storing them.  tests/philox_ref.py is the numpy restatement.
#pragma once
#include <stdint.h>

namespace philox {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

struct Dropout {
  uint2 key;          // 64-bit seed
  uint32_t thresh;    // an element is dropped when its 16-bit random number is < thresh  (p = thresh / 65536)
  float inv_keep;     // 1 / (1 - p)
};

// keep-mask bits of the 16 consecutive keys [key16*16, key16*16 + 16) of packed attention row `row`:
// bit c set = element (row, key16*16 + c) is kept.  Two Philox calls, eight 16-bit numbers each.
__device__ __forceinline__ uint32_t keep_mask16(const Dropout& d, uint32_t row, uint32_t key16) {
  uint32_t m = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint4 r = philox4x32_10(make_uint4(key16 * 2u + half, row, 0u, 0u), d.key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m |= ((w[j] & 0xFFFFu) >= d.thresh ? 1u : 0u) << (half * 8 + 2 * j);
      m |= ((w[j] >> 16) >= d.thresh ? 1u : 0u) << (half * 8 + 2 * j + 1);
    }
  }
  return m;
}

__host__ __device__ inline uint32_t thresh_of(float p) {
  float t = p * 65536.0f + 0.5f;
  if (t < 0.f) t = 0.f;
  if (t > 65535.f) t = 65535.f;
  return (uint32_t)t;
}

}  // namespace philox
