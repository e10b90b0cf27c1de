// Shared device helpers for the sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace srb {

constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ull;

__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t h) {
  h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 33;
  return h;
}

// The four table rows of one attribute id (must match TorchOps.hash_rows).
__device__ __forceinline__ void hash_rows(uint64_t id, uint64_t seed, uint32_t n_rows, uint32_t rows[4]) {
  uint64_t h1 = fmix64(id ^ (seed * kGolden));
  uint64_t h2 = fmix64(h1 + kGolden);
  rows[0] = (uint32_t)(h1 & 0xFFFFFFFFull) % n_rows;
  rows[1] = (uint32_t)(h1 >> 32) % n_rows;
  rows[2] = (uint32_t)(h2 & 0xFFFFFFFFull) % n_rows;
  rows[3] = (uint32_t)(h2 >> 32) % n_rows;
}

// Counter-based dropout (must match TorchOps.dropout_mask): keep iff u >= p.
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  uint64_t h = fmix64(idx + seed * kGolden);
  float u = (float)(h >> 40) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.0f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

struct alignas(16) bf16x8 { __nv_bfloat16 v[8]; };

}  // namespace srb
