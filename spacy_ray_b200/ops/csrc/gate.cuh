// System-scope flag primitives + the consumer-side "published" gate.
//
// C2: the parameter pull of the reference (owner pushes each updated
// tensor to every peer, proxies.py:71-75; the peer adopts it at its next read,
// proxies.py:111-118) becomes: the owner's exchange kernel stores its refreshed bf16 shard
// straight into every rank's weight buffer and then bumps a per-(bucket, owner) epoch flag in
// every rank's signal page.  Nobody waits for that at the end of the step.  Instead the FIRST
// kernel of the next forward pass that reads a bucket's weights - the tcgen05 GEMM's TMA
// producer warp (gemm_tcgen05.cu) or hash_embed_fwd_kernel - spins on those flags
// (ld.acquire.sys) right before its first weight load, so the all-gather's tail overlaps
// the consumer's launch, prologue (barrier init, TMEM allocation, descriptor prefetch) and the
// kernels of earlier buckets.
#pragma once
#include <stdint.h>

#include "comm_launch.h"

namespace srb {

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy writes (peers' stores observed through the acquire above) -> async-proxy reads (TMA)
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// Spin until *flag >= target (system scope).  Returns false on timeout.
__device__ __forceinline__ bool wait_flag_sys(const uint32_t* flag, uint32_t target, uint64_t timeout_ns) {
  if ((int32_t)(ld_acquire_sys(flag) - target) >= 0) return true;
  const uint64_t t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(flag) - target) < 0) {
    if (globaltimer_ns() - t0 > timeout_ns) return false;
    __nanosleep(32);
  }
  return true;
}

// Called by ALL 32 lanes of one warp.  Lane r polls owner r's flag of each bucket in the mask.
// On timeout the error word is set to 7 and the caller proceeds (the host raises at its next check).
__device__ __forceinline__ void gate_wait_warp(const GateArgs& g) {
  if (g.flags == nullptr || g.mask == 0u) return;
  const int lane = threadIdx.x & 31;
  if (lane < g.world) {
    const uint32_t e = *(const volatile uint32_t*)g.epoch;
    uint32_t m = g.mask;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (!wait_flag_sys(g.flags + flag_pub_idx(b, lane), e, g.timeout_ns)) {
        atomicExch(g.error, 7);
        break;
      }
    }
  }
  __syncwarp();
}

}  // namespace srb
