// Philox4x32-10 counter-based random numbers (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), device side.
// Replaces the torch.rand / torch.randn calls on the step path (sample_time_interval motion_lib_base.py:412, the action noise of
// ModelA2CContinuousLogStd [rl_games], the reparameterisation noise amp_network_z_builder.py:243-246) with draws made inside the
// consuming kernel: stream = (seed, offset), element = (row, lane group).  Deterministic for a given (seed, offset, index).
#pragma once
#include <stdint.h>

namespace pulse {

struct Philox4 {
  unsigned x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(unsigned long long seed, unsigned long long index, unsigned long long offset) {
  unsigned k0 = static_cast<unsigned>(seed), k1 = static_cast<unsigned>(seed >> 32);
  unsigned c0 = static_cast<unsigned>(index), c1 = static_cast<unsigned>(index >> 32);
  unsigned c2 = static_cast<unsigned>(offset), c3 = static_cast<unsigned>(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return {c0, c1, c2, c3};
}

// uniform in [0, 1) on the 2^-24 grid (what torch.rand produces for float32)
__device__ __forceinline__ float u01(unsigned x) { return static_cast<float>(x >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long index, unsigned long long offset) {
  return u01(philox4x32_10(seed, index, offset).x);
}

// two independent standard normals from two 32-bit words (Box-Muller; u1 in (0, 1] so the logarithm is finite)
__device__ __forceinline__ void box_muller(unsigned a, unsigned b, float& n0, float& n1) {
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const float u2 = static_cast<float>(b >> 8) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.28318530717958647692f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

}  // namespace pulse
