// Shared device helpers for the loner_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/loner_b200.h"

#define LONER_CHECK_LAUNCH()                                   \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) return LONER_E_LAUNCH;             \
  } while (0)

namespace loner {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// inclusive scan (sum) across the warp
__device__ __forceinline__ float warp_incl_scan_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// inclusive scan (product) across the warp
__device__ __forceinline__ float warp_incl_scan_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}

// inclusive suffix scan (sum): result = sum over lanes >= this lane
__device__ __forceinline__ float warp_incl_suffix_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

// ---- Philox4x32-10 counter-based generator (Salmon et al. 2011), used when the caller does
// not inject the random numbers of ray_sampling.py:71-72 / rendering_tcnn.py:48,104.
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint32_t stream) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x9E3779B9u;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

__device__ __forceinline__ float u32_to_unit(uint32_t x) {  // [0,1), 24 bits like torch.rand
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);   // (0,1]
  float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// torch.linspace(0,1,H)[i] exactly as ATen computes it (symmetric halves).
__device__ __forceinline__ float linspace01(int i, int H) {
  float step = 1.0f / (float)(H - 1);
  return (i < H / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(H - 1 - i)));
}

}  // namespace loner
