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

// Counter-based dropout (must match TorchOps.dropout_mask).  Elements are hashed in groups of
// four: one fmix64 per group, 16 bits per element; keep iff u16 >= thr, thr = floor(p * 2^16).
__host__ __device__ __forceinline__ uint32_t dropout_thr(float p) { return (uint32_t)(p * 65536.0f); }
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  const uint64_t h = fmix64((idx >> 2) + seed * kGolden);
  const uint32_t u = (uint32_t)(h >> (16 * (uint32_t)(idx & 3))) & 0xFFFFu;
  return u >= dropout_thr(p) ? inv_keep : 0.0f;
}
// Eight consecutive elements starting at idx (idx % 8 == 0): two hashes.
__device__ __forceinline__ void dropout_scale8(uint64_t seed, uint64_t idx, uint32_t thr, float inv_keep,
                                               float out[8]) {
  const uint64_t base = (idx >> 2) + seed * kGolden;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const uint64_t h = fmix64(base + g);
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    out[g * 4 + 0] = (lo & 0xFFFFu) >= thr ? inv_keep : 0.f;
    out[g * 4 + 1] = (lo >> 16) >= thr ? inv_keep : 0.f;
    out[g * 4 + 2] = (hi & 0xFFFFu) >= thr ? inv_keep : 0.f;
    out[g * 4 + 3] = (hi >> 16) >= thr ? inv_keep : 0.f;
  }
}

// Programmatic dependent launch (see launch.h): let the next kernel of the stream be scheduled
// now, and hold this one's first global access until its predecessor has completed and flushed.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }

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
