// C1 + K8 + C2 in ONE kernel: reduce-scatter of the flat gradient bucket over NVLink
// peer memory, fused with the sharded Adam step (per-tensor clipping, fp32 master) and
// the all-gather push of the refreshed bf16 weights into every peer's weight buffer.
//
// This replaces the reference's per-key Ray RPCs (gradient push proxies.py:104,
// parameter push proxies.py:75): same bytes, zero host involvement, no NCCL call.
//
//   phase 0  flag exchange: "my gradients for this epoch are complete"
//   phase 1  for my shard: g = sum_p peer_grad_p[shard]  (P2P ld.relaxed.sys v4, or one
//            multimem.ld_reduce through the NVSwitch when a multicast mapping exists)
//            -> written in place + per-key sum of squares
//   grid barrier
//   phase 2  per-key clip, Adam on fp32 master/m1/m2, bf16 weights stored straight into
//            all W ranks' weight buffers (P2P st v4 / multimem.st), own grads zeroed
//   phase 3  flag exchange "read done" -> zero the rest of my gradient buffer;
//            flag exchange "weights of epoch e published" -> (optionally) wait for all.
//
// Every spin has a wall-clock timeout (%globaltimer); on timeout the kernel records an
// error code and exits instead of hanging the GPU (a dead peer must
// produce an error, not a hang).
#include "comm_launch.h"
#include "common.cuh"

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
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v2_b32(void* mc, uint32_t a, uint32_t b) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(a)),
               "f"(__uint_as_float(b))
               : "memory");
}

// Spin until *flag >= target (system scope).  Returns false on timeout.
__device__ __forceinline__ bool wait_flag_sys(const uint32_t* flag, uint32_t target, uint64_t timeout_ns) {
  const uint64_t t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(flag) - target) < 0) {
    if (globaltimer_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
  return true;
}

// Device-scope barrier across the (co-resident) CTAs of this kernel.
__device__ __forceinline__ bool grid_barrier(uint32_t* counter, uint32_t target, uint64_t timeout_ns) {
  __syncthreads();
  bool ok = true;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_gpu(counter) - target) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) { ok = false; break; }
    }
  }
  __syncthreads();
  return ok;
}

__global__ void __launch_bounds__(256, 2) fused_rs_adam_ag_kernel(FusedCommArgs a) {
  const int W = a.world, rank = a.rank;
  const uint32_t epoch = *a.epoch + 1;                     // flag value for this invocation
  const uint64_t tmo = a.timeout_ns;
  __shared__ int s_fail;
  __shared__ float s_part[8];
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();

  // ---------------- phase 0: gradients of every rank are complete -----------------------
  if (blockIdx.x == 0 && threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.signal[threadIdx.x] + kSlotGrad * kMaxWorld + rank, epoch);
  }
  if (threadIdx.x < W) {
    if (!wait_flag_sys(a.signal[rank] + kSlotGrad * kMaxWorld + threadIdx.x, epoch, tmo)) s_fail = 1;
  }
  __syncthreads();
  if (s_fail) { if (threadIdx.x == 0) atomicExch(a.error, 1); return; }

  const int64_t s0 = a.shard_start;
  float* my_grad = a.grad[rank];
  const float gs = a.hyper[7];
  // ---------------- phase 1: reduce my shard, per-key sum of squares ---------------------
  for (int b = blockIdx.x; b < a.n_blocks; b += gridDim.x) {
    const int k = a.blk_key[b];
    const int64_t base = a.key_off[k] + (int64_t)a.blk_off[b] * kCommChunk;
    const int64_t end = a.key_off[k] + a.key_len[k];
    float acc = 0.f;
    {
      // 4 independent 16-byte loads per peer are issued before any is consumed (the work item
      // is exactly 256 threads x 4 vectors), so each thread keeps >= 4 NVLink reads in flight.
      float4 s[4];
      bool ok[4];
      const int64_t t0 = base + threadIdx.x * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) { ok[j] = (t0 + j * 1024) < end; s[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
      if (a.grad_mc) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (ok[j]) s[j] = multimem_ld_reduce_v4(a.grad_mc + s0 + t0 + j * 1024);
      } else {
#pragma unroll 2
        for (int p = 0; p < W; ++p) {
          const float* src = a.grad[(rank + p) % W] + s0 + t0;      // stagger peers across ranks
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) if (ok[j]) v[j] = ld_relaxed_sys_v4(src + j * 1024);
#pragma unroll
          for (int j = 0; j < 4; ++j) if (ok[j]) { s[j].x += v[j].x; s[j].y += v[j].y; s[j].z += v[j].z; s[j].w += v[j].w; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ok[j]) {
          s[j].x *= gs; s[j].y *= gs; s[j].z *= gs; s[j].w *= gs;
          *(float4*)(my_grad + s0 + t0 + j * 1024) = s[j];
          acc += s[j].x * s[j].x + s[j].y * s[j].y + s[j].z * s[j].z + s[j].w * s[j].w;
        }
      }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += s_part[i];
      atomicAdd(a.norms_sq + k, t);
    }
    __syncthreads();
  }
  // everyone has finished READING peers' gradient buffers -> they may be zeroed
  const uint32_t bar_base = (epoch - 1) * 3u * gridDim.x;
  if (!grid_barrier(a.bar_counter, bar_base + gridDim.x, tmo)) { if (threadIdx.x == 0) atomicExch(a.error, 2); return; }
  if (blockIdx.x == 0 && threadIdx.x < W)
    st_release_sys(a.signal[threadIdx.x] + kSlotRead * kMaxWorld + rank, epoch);

  // ---------------- phase 2: clip + Adam + publish bf16 weights to all ranks -------------
  const float lr = a.hyper[0], b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], clip = a.hyper[4], l2 = a.hyper[5];
  const bool wd = a.hyper[6] != 0.f;
  const float t = (float)(*a.step + 1);
  const float fix1 = 1.f - powf(b1, t), fix2 = 1.f - powf(b2, t);
  const float lr_t = lr * sqrtf(fix2) / fix1;
  for (int b = blockIdx.x; b < a.n_blocks; b += gridDim.x) {
    const int k = a.blk_key[b];
    const int64_t base = a.key_off[k] + (int64_t)a.blk_off[b] * kCommChunk;
    const int64_t end = a.key_off[k] + a.key_len[k];
    float scale = 1.f;
    if (clip > 0.f) {
      const float norm = sqrtf(a.norms_sq[k]);
      if (norm >= clip) scale = clip / fmaxf(norm, 1e-30f);
    }
    {
      // loads for all 4 vectors of this thread first (16 independent 16-byte loads in flight)
      const int64_t t0 = base + threadIdx.x * 4;
      float4 g4[4], w4[4], a4[4], b4[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t idx = t0 + j * 1024;
        ok[j] = idx < end;
        if (ok[j]) {
          g4[j] = *(const float4*)(my_grad + s0 + idx);
          w4[j] = *(const float4*)(a.master + idx);
          a4[j] = *(const float4*)(a.m1 + idx);
          b4[j] = *(const float4*)(a.m2 + idx);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!ok[j]) continue;
        const int64_t idx = t0 + j * 1024;
        float gv[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w}, wv[4] = {w4[j].x, w4[j].y, w4[j].z, w4[j].w};
        float av[4] = {a4[j].x, a4[j].y, a4[j].z, a4[j].w}, bv[4] = {b4[j].x, b4[j].y, b4[j].z, b4[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = gv[e];
          if (l2 != 0.f && !wd) x += l2 * wv[e];
          x *= scale;
          av[e] = b1 * av[e] + (1.f - b1) * x;
          bv[e] = b2 * bv[e] + (1.f - b2) * x * x;
          wv[e] -= lr_t * av[e] / (sqrtf(bv[e]) + eps);
          if (wd && l2 != 0.f) wv[e] *= (1.f - lr * l2);
        }
        *(float4*)(a.master + idx) = make_float4(wv[0], wv[1], wv[2], wv[3]);
        *(float4*)(a.m1 + idx) = make_float4(av[0], av[1], av[2], av[3]);
        *(float4*)(a.m2 + idx) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        *(float4*)(my_grad + s0 + idx) = make_float4(0.f, 0.f, 0.f, 0.f);
        __nv_bfloat162 lo = __floats2bfloat162_rn(wv[0], wv[1]);
        __nv_bfloat162 hi = __floats2bfloat162_rn(wv[2], wv[3]);
        const uint32_t ulo = *(uint32_t*)&lo, uhi = *(uint32_t*)&hi;
        if (a.param_mc) {
          multimem_st_v2_b32((__nv_bfloat16*)a.param_mc + s0 + idx, ulo, uhi);
        } else {
#pragma unroll 8
          for (int p = 0; p < W; ++p) {
            const int peer = (rank + p) % W;
            *(uint2*)((__nv_bfloat16*)a.param[peer] + s0 + idx) = make_uint2(ulo, uhi);
          }
        }
      }
    }
  }
  __threadfence_system();
  if (!grid_barrier(a.bar_counter, bar_base + 2u * gridDim.x, tmo)) { if (threadIdx.x == 0) atomicExch(a.error, 3); return; }
  if (blockIdx.x == 0 && threadIdx.x < W)
    st_release_sys(a.signal[threadIdx.x] + kSlotParam * kMaxWorld + rank, epoch);

  // ---------------- phase 3: zero the peers' shards of my gradient buffer ----------------
  if (threadIdx.x < W) {
    if (!wait_flag_sys(a.signal[rank] + kSlotRead * kMaxWorld + threadIdx.x, epoch, tmo)) s_fail = 1;
  }
  __syncthreads();
  if (s_fail) { if (threadIdx.x == 0) atomicExch(a.error, 4); return; }
  {
    const int64_t total4 = a.total_elems / 4;
    const int64_t lo4 = s0 / 4, hi4 = (s0 + a.shard_cap) / 4;
    float4* g4 = (float4*)my_grad;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < total4; i += (int64_t)gridDim.x * 256)
      if (i < lo4 || i >= hi4) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // padding of my own shard beyond the owned keys is never written, stays zero
  }
  for (int k = blockIdx.x * 256 + threadIdx.x; k < a.n_keys; k += gridDim.x * 256) a.norms_sq[k] = 0.f;
  // ---------------- all weights of this epoch have landed here ---------------------------
  if (a.wait_params) {
    if (threadIdx.x < W) {
      if (!wait_flag_sys(a.signal[rank] + kSlotParam * kMaxWorld + threadIdx.x, epoch, tmo)) s_fail = 1;
    }
    __syncthreads();
    if (s_fail) { if (threadIdx.x == 0) atomicExch(a.error, 5); return; }
  }
  if (!grid_barrier(a.bar_counter, bar_base + 3u * gridDim.x, tmo)) { if (threadIdx.x == 0) atomicExch(a.error, 6); return; }
  if (blockIdx.x == 0 && threadIdx.x == 0) { *a.epoch = epoch; *a.step = *a.step + 1; }
}

cudaError_t launch_fused_rs_adam_ag(const FusedCommArgs& a, int grid, cudaStream_t s) {
  fused_rs_adam_ag_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------
// Stand-alone collectives on the same machinery (bandwidth sweep, BASELINE.json config 5)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) p2p_reduce_scatter_kernel(P2PCollArgs a) {
  const int W = a.world, rank = a.rank;
  const uint32_t epoch = *a.epoch + 1;
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.signal[threadIdx.x] + kSlotGrad * kMaxWorld + rank, epoch);
  }
  if (threadIdx.x < W && !wait_flag_sys(a.signal[rank] + kSlotGrad * kMaxWorld + threadIdx.x, epoch, a.timeout_ns)) s_fail = 1;
  __syncthreads();
  if (s_fail) { if (threadIdx.x == 0) atomicExch(a.error, 1); return; }
  const int64_t n4 = a.shard_elems / 4;
  const float* const* src = (const float* const*)a.buf;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i0 = blockIdx.x * 256 + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    float4 acc[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { ok[j] = (i0 + j * stride) < n4; acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    if (a.mc) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ok[j]) acc[j] = multimem_ld_reduce_v4((const float*)a.mc + a.shard_elems * rank + (i0 + j * stride) * 4);
    } else {
#pragma unroll 2
      for (int p = 0; p < W; ++p) {
        const float* sp = src[(rank + p) % W] + a.shard_elems * rank;
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) if (ok[j]) v[j] = ld_relaxed_sys_v4(sp + (i0 + j * stride) * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (ok[j]) { acc[j].x += v[j].x; acc[j].y += v[j].y; acc[j].z += v[j].z; acc[j].w += v[j].w; }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) if (ok[j]) *(float4*)((float*)a.out + (i0 + j * stride) * 4) = acc[j];
  }
  const uint32_t bar_base = (epoch - 1) * gridDim.x;
  if (!grid_barrier(a.bar_counter, bar_base + gridDim.x, a.timeout_ns)) { if (threadIdx.x == 0) atomicExch(a.error, 2); return; }
  // peers may only overwrite their buffers once everyone has read: exchange read-done flags
  if (blockIdx.x == 0 && threadIdx.x < W)
    st_release_sys(a.signal[threadIdx.x] + kSlotRead * kMaxWorld + rank, epoch);
  if (blockIdx.x == 0) {
    if (threadIdx.x < W && !wait_flag_sys(a.signal[rank] + kSlotRead * kMaxWorld + threadIdx.x, epoch, a.timeout_ns)) s_fail = 1;
    __syncthreads();
    if (threadIdx.x == 0) { if (s_fail) atomicExch(a.error, 3); *a.epoch = epoch; }
  }
}

__global__ void __launch_bounds__(256, 1) p2p_all_gather_kernel(P2PCollArgs a) {
  const int W = a.world, rank = a.rank;
  const uint32_t epoch = *a.epoch + 1;
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  // my shard (a.out, shard_elems fp32-sized units of 4 bytes) -> every rank's buf[rank*shard ...]
  const int64_t n4 = a.shard_elems / 4;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = *(const float4*)((const float*)a.out + i * 4);
    const int64_t idx = a.shard_elems * rank + i * 4;
    if (a.mc) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"((float*)a.mc + idx), "f"(v.x),
                   "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    } else {
#pragma unroll 8
      for (int p = 0; p < W; ++p) *(float4*)((float*)a.buf[(rank + p) % W] + idx) = v;
    }
  }
  __threadfence_system();
  const uint32_t bar_base = (epoch - 1) * gridDim.x;
  if (!grid_barrier(a.bar_counter, bar_base + gridDim.x, a.timeout_ns)) { if (threadIdx.x == 0) atomicExch(a.error, 2); return; }
  if (blockIdx.x == 0) {
    if (threadIdx.x < W) st_release_sys(a.signal[threadIdx.x] + kSlotParam * kMaxWorld + rank, epoch);
    if (threadIdx.x < W && !wait_flag_sys(a.signal[rank] + kSlotParam * kMaxWorld + threadIdx.x, epoch, a.timeout_ns)) s_fail = 1;
    __syncthreads();
    if (threadIdx.x == 0) { if (s_fail) atomicExch(a.error, 3); *a.epoch = epoch; }
  }
}

cudaError_t launch_p2p_reduce_scatter(const P2PCollArgs& a, int grid, cudaStream_t s) {
  p2p_reduce_scatter_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_p2p_all_gather(const P2PCollArgs& a, int grid, cudaStream_t s) {
  p2p_all_gather_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace srb
