// C1 + K8 + C2: the gradient push / parameter pull of the reference as peer-memory kernels.
//
// The reference ships every gradient tensor to its owner and every updated tensor back to
// every peer as separate Ray RPCs (gradient push proxies.py:102-104, optimizer on the owner
// proxies.py:126-128, parameter push proxies.py:71-75).  Here the flat gradient bucket is cut
// into a few BUCKETS in the order the backward pass completes them, and for every bucket the
// kernel pair `bucket_reduce_kernel` / `bucket_update_kernel` runs - on a side stream, concurrently
// with the rest of the backward pass - the whole owner-side pipeline over NVLink peer memory,
// no NCCL, no host:
//
//   phase 0  "my gradients of bucket b, epoch e, are complete" -> every rank's signal page;
//            owners wait for all ranks' flags
//   phase 1  for my keys of the bucket: g = sum_p peer_grad_p  (one multimem.ld_reduce through the
//            NVSwitch, or P2P ld.relaxed.sys from every peer), g -> local scratch, per-key sum of squares
//   (kernel boundary: one CTA per 4096-element work item, so the two phases are two launches)
//   phase 2  per-key clip, Adam / RAdam / SGD on the fp32 master (+ moments, + parameter averages),
//            bf16 weights stored straight into ALL ranks' weight buffers (multimem.st / P2P st);
//            the last CTA to finish releases "published (b, me) = e" into every rank's signal page.
//
// Nothing waits for the publication here: the first consumer of a bucket's weights in the next
// forward pass does (gate.cuh; gemm_tcgen05.cu's TMA producer warp, hash_embed_fwd_kernel).
// The same flag tells a rank that the owners are done READING its gradients of the bucket: the
// accumulators are cleared locally at the start of the next step (bucket_gate_zero_kernel, side
// stream, under the forward pass) - no clear stores over NVLink, no "read done" flag round.
//
// Every spin has a wall-clock timeout (%globaltimer); on timeout the kernel records an
// error code and exits instead of hanging the GPU (a dead peer must
// produce an error, not a hang).
#include "comm_launch.h"
#include "common.cuh"
#include "gate.cuh"

namespace srb {

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
__device__ __forceinline__ void multimem_st_v4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st_v4_b32(void* mc, const uint32_t v[4]) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(v[0])),
               "f"(__uint_as_float(v[1])), "f"(__uint_as_float(v[2])), "f"(__uint_as_float(v[3]))
               : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// One work item (= one 4096-element chunk of one owned key) per CTA, two launches per bucket:
//   bucket_signal_kernel / bucket_wait_kernel  phase 0 (one warp each: my grad-ready flag out; everybody's in)
//   bucket_reduce_kernel  phase 1 (reduce + clear, per-key sum of squares)
//   bucket_update_kernel  phase 2 (clip, optimizer, weight stores) + publication by the last CTA
// The kernel boundary is the "all of this key's partial norms are in" barrier.  No device-wide
// barrier inside a kernel means no co-residency requirement: the grids can be as large as the
// bucket, and the CTAs (80 registers x 256 threads) slot in next to the CTAs of whatever the
// backward pass is running (a tcgen05 GEMM CTA leaves room for exactly one of them per SM).
// phase 0 as ONE-WARP kernels: "my gradients of bucket b are complete" -> every rank's signal page
// (bucket_signal_kernel, on its own stream: it must go out the moment the gradients exist, not after
// the exchange of the previous bucket), and the wait for everybody else's (bucket_wait_kernel, at the
// head of the bucket's exchange).  The wait can last as long as the slowest rank's backward pass;
// doing it inside the reduce kernel would park hundreds of spinning CTAs on the SMs and starve this
// rank's own backward GEMMs of registers exactly while they should be running.
__global__ void __launch_bounds__(32) bucket_signal_kernel(FusedCommArgs a) {
  const int W = a.world, rank = a.rank, bkt = a.bucket;
  const uint32_t epoch = *(const volatile uint32_t*)a.epoch + 1;   // flag value of this exchange
  const int lane = threadIdx.x;
  unsigned long long* tr = a.trace ? a.trace + (size_t)bkt * kTraceWords : nullptr;
  if (tr && lane == 0) {
    tr[0] = globaltimer_ns(); tr[1] = 0ull; tr[2] = ~0ull; tr[3] = 0ull; tr[4] = ~0ull; tr[5] = 0ull;
  }
  if (W > 1 && lane < W) {
    fence_acq_rel_sys();
    st_release_sys(a.signal[lane] + flag_grad_idx(bkt, rank), epoch);
  }
}

__global__ void __launch_bounds__(32) bucket_wait_kernel(FusedCommArgs a) {
  const int W = a.world, rank = a.rank, bkt = a.bucket;
  const uint32_t epoch = *(const volatile uint32_t*)a.epoch + 1;
  const int lane = threadIdx.x;
  unsigned long long* tr = a.trace ? a.trace + (size_t)bkt * kTraceWords : nullptr;
  if (a.blk_end - a.blk_begin == 0) {
    // nothing of this bucket is mine: there is nothing to wait for, and nothing to publish but the flag
    // (consumers wait for every rank's flag of a bucket)
    if (W > 1 && lane < W) st_release_sys(a.signal[lane] + flag_pub_idx(bkt, rank), epoch);
    if (a.last && lane == 0) {             // ... except, for the step's last bucket, the counters
      *a.step = *(const volatile int32_t*)a.step + 1;
      __threadfence();
      *(volatile uint32_t*)a.epoch = epoch;
    }
    return;
  }
  if (W > 1 && lane < W && !wait_flag_sys(a.signal[rank] + flag_grad_idx(bkt, lane), epoch, a.timeout_ns))
    atomicExch(a.error, 1);
  __syncwarp();
  if (tr && lane == 0) tr[1] = globaltimer_ns();
}

// Clear my gradient accumulators of a bucket once every owner has published it (= finished reading
// them).  Runs at the start of the next step on a side stream, under the forward pass; the first
// gradient write of the step waits for it.  (Round 1 zero-filled inside the exchange kernel after a
// "read done" flag round; an owner-side multimem.st of zeros doubled the NVLink bytes of the reduce.)
__global__ void __launch_bounds__(256) bucket_gate_zero_kernel(GateArgs g, float* __restrict__ grad,
                                                               const int64_t* __restrict__ ext_off,
                                                               const int64_t* __restrict__ ext_len, int ext_begin,
                                                               int ext_end) {
  if (threadIdx.x < 32) gate_wait_warp(g);
  __syncthreads();
  for (int e = ext_begin; e < ext_end; ++e) {
    float4* p = (float4*)(grad + ext_off[e]);
    const int64_t n4 = ext_len[e] >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
      p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(256, 3) bucket_reduce_kernel(FusedCommArgs a) {
  const int W = a.world, rank = a.rank;
  const int n_items = a.blk_end - a.blk_begin;
  __shared__ float s_part[8];
  if (n_items == 0 || *(const volatile int32_t*)a.error != 0) return;
  unsigned long long* tr = a.trace ? a.trace + (size_t)a.bucket * kTraceWords : nullptr;
  if (tr && threadIdx.x == 0) atomicMin(tr + 2, (unsigned long long)globaltimer_ns());
  const int64_t s0 = a.shard_start;
  float* my_grad = a.grad[rank];
  const float l2 = a.hyper[5];
  const bool wd = a.hyper[6] != 0.f;
  const float gs = a.hyper[7];
  const bool l2_in_grad = (l2 != 0.f) && !wd;
  const bool keep_red = (W > 1) || gs != 1.f || l2_in_grad;       // otherwise the update re-reads the gradient itself
  // ---------------- phase 1: reduce my keys, clear every copy, per-key sum of squares --------
  const int b = a.blk_begin + blockIdx.x;
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
    if (W == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (ok[j]) s[j] = *(const float4*)(my_grad + s0 + t0 + j * 1024);
    } else if (a.grad_mc) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (ok[j]) s[j] = multimem_ld_reduce_v4(a.grad_mc + s0 + t0 + j * 1024);
    } else {
#pragma unroll 2
      for (int p = 0; p < W; ++p) {
        float* src = a.grad[(rank + p) % W] + s0 + t0;            // stagger peers across ranks
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) if (ok[j]) v[j] = ld_relaxed_sys_v4(src + j * 1024);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (ok[j]) {
          s[j].x += v[j].x; s[j].y += v[j].y; s[j].z += v[j].z; s[j].w += v[j].w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (ok[j]) {
        const int64_t idx = t0 + j * 1024;
        s[j].x *= gs; s[j].y *= gs; s[j].z *= gs; s[j].w *= gs;
        if (l2_in_grad) {                       // thinc: L2 joins the gradient BEFORE the clip norm
          const float4 w4 = *(const float4*)(a.master + idx);
          s[j].x += l2 * w4.x; s[j].y += l2 * w4.y; s[j].z += l2 * w4.z; s[j].w += l2 * w4.w;
        }
        if (keep_red) *(float4*)(a.red + idx) = s[j];
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
    if (tr) atomicMax(tr + 3, (unsigned long long)globaltimer_ns());
  }
}

__global__ void __launch_bounds__(256, 3) bucket_update_kernel(FusedCommArgs a) {
  const int W = a.world, rank = a.rank, bkt = a.bucket;
  const uint32_t epoch = *(const volatile uint32_t*)a.epoch + 1;
  const int n_items = a.blk_end - a.blk_begin;
  __shared__ int s_last;
  unsigned long long* tr = a.trace ? a.trace + (size_t)bkt * kTraceWords : nullptr;
  if (tr && threadIdx.x == 0) atomicMin(tr + 4, (unsigned long long)globaltimer_ns());
  if (n_items > 0 && *(const volatile int32_t*)a.error == 0) {
    if (a.test_delay_ns) {          // test hook: a slow owner (the consumers' gates must hold them back)
      if (threadIdx.x == 0) {
        const uint64_t t0 = globaltimer_ns();
        while (globaltimer_ns() - t0 < a.test_delay_ns) __nanosleep(1000);
      }
      __syncthreads();
    }
    const int64_t s0 = a.shard_start;
    float* my_grad = a.grad[rank];
    const float lr = a.hyper[0], b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], clip = a.hyper[4], l2 = a.hyper[5];
    const bool wd = a.hyper[6] != 0.f;
    const float gs = a.hyper[7];
    const bool keep_red = (W > 1) || gs != 1.f || ((l2 != 0.f) && !wd);
    // ---------------- phase 2: clip + optimizer + publish bf16 weights to all ranks ----------
    const float t = (float)(*(const volatile int32_t*)a.step + 1);
    const float b1t = powf(b1, t), b2t = powf(b2, t);
    const float fix1 = 1.f - b1t, fix2 = 1.f - b2t;
    float lr_t = lr * sqrtf(fix2) / fix1;                    // Adam: bias correction folded into the rate
    bool rect = true;
    if (a.opt_mode == kOptRAdam) {
      const float sma_max = 2.f / (1.f - b2) - 1.f;
      const float sma = sma_max - 2.f * t * b2t / fix2;
      rect = sma >= 5.f;
      lr_t = rect ? lr * sqrtf(fix2 * (sma - 4.f) / (sma_max - 4.f) * (sma - 2.f) / sma * sma_max / (sma_max - 2.f)) / fix1
                  : lr / fix1;
    }
    // thinc's update_averages: decay = min((1 + t) / (10 + t), 0.9999); ema -= (1 - decay) * (ema - w)
    const float avg_mix = 1.f - fminf((1.f + t) / (10.f + t), 0.9999f);
    const int b = a.blk_begin + blockIdx.x;
    const int k = a.blk_key[b];
    const int64_t base = a.key_off[k] + (int64_t)a.blk_off[b] * kCommChunk;
    const int64_t end = a.key_off[k] + a.key_len[k];
    float scale = 1.f;
    if (clip > 0.f) {
      const float norm = sqrtf(a.norms_sq[k]);
      if (norm >= clip) scale = clip / fmaxf(norm, 1e-30f);
    }
    // Two passes; in each a thread owns EIGHT consecutive elements: 2 x 5 independent 16-byte loads in
    // flight, and the refreshed weights leave as ONE 16-byte bf16x8 store per thread and pass (8-byte
    // multimem / peer stores were the slow part of this kernel with peers: half the payload per
    // NVLink packet).  80 registers: three CTAs fit on an SM - or one NEXT TO a tcgen05 GEMM CTA.
    const float* gsrc = keep_red ? a.red : (my_grad + s0);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int64_t idx = base + half * 2048 + threadIdx.x * 8;
      if (idx >= end) continue;                       // key_len is a multiple of 128: all 8 or none
      float4 g4[2], w4[2], a4[2], b4[2], e4[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        g4[j] = *(const float4*)(gsrc + idx + 4 * j);
        w4[j] = *(const float4*)(a.master + idx + 4 * j);
        if (a.opt_mode != kOptSGD) {
          a4[j] = *(const float4*)(a.m1 + idx + 4 * j);
          b4[j] = *(const float4*)(a.m2 + idx + 4 * j);
        }
        if (a.avg) e4[j] = *(const float4*)(a.avg + idx + 4 * j);
      }
      uint32_t packed[4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float gv[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w}, wv[4] = {w4[j].x, w4[j].y, w4[j].z, w4[j].w};
        float av[4] = {a4[j].x, a4[j].y, a4[j].z, a4[j].w}, bv[4] = {b4[j].x, b4[j].y, b4[j].z, b4[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x = gv[e] * scale;
          if (a.opt_mode == kOptSGD) {
            wv[e] -= lr * x;
          } else {
            av[e] = b1 * av[e] + (1.f - b1) * x;
            bv[e] = b2 * bv[e] + (1.f - b2) * x * x;
            if (rect) wv[e] -= lr_t * av[e] / (sqrtf(bv[e]) + eps);
            else wv[e] -= lr_t * av[e];
          }
          if (wd && l2 != 0.f) wv[e] *= (1.f - lr * l2);
        }
        *(float4*)(a.master + idx + 4 * j) = make_float4(wv[0], wv[1], wv[2], wv[3]);
        if (a.opt_mode != kOptSGD) {
          *(float4*)(a.m1 + idx + 4 * j) = make_float4(av[0], av[1], av[2], av[3]);
          *(float4*)(a.m2 + idx + 4 * j) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        }
        if (a.avg) {
          float ev[4] = {e4[j].x, e4[j].y, e4[j].z, e4[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) ev[e] -= avg_mix * (ev[e] - wv[e]);
          *(float4*)(a.avg + idx + 4 * j) = make_float4(ev[0], ev[1], ev[2], ev[3]);
        }
        if (W == 1) *(float4*)(my_grad + s0 + idx + 4 * j) = make_float4(0.f, 0.f, 0.f, 0.f);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(wv[0], wv[1]);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(wv[2], wv[3]);
        packed[2 * j] = *(const uint32_t*)&lo;
        packed[2 * j + 1] = *(const uint32_t*)&hi;
      }
      if (a.param_mc) {
        multimem_st_v4_b32((__nv_bfloat16*)a.param_mc + s0 + idx, packed);
      } else {
#pragma unroll 8
        for (int p = 0; p < W; ++p) {
          const int peer = (rank + p) % W;
          *(uint4*)((__nv_bfloat16*)a.param[peer] + s0 + idx) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
      }
    }
  }
  // ---------------- the last CTA to finish publishes the bucket --------------------------------
  // (stores -> fence -> counter: the classic last-block pattern; the publishing CTA's acquire of the
  //  counter + its release of the flag order every CTA's weight / zero stores before the flag)
  __syncthreads();                        // every thread's stores are ordered before thread 0's fence (CTA scope) ...
  if (threadIdx.x == 0) {
    if (W > 1) fence_acq_rel_sys();       // ... which releases them at system scope: ONE fence per CTA, not 256
    else fence_acq_rel_gpu();
    const uint32_t prev = atomicAdd(a.bar + 2 * bkt, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (W > 1) fence_acq_rel_sys(); else fence_acq_rel_gpu();
  if (threadIdx.x == 0) a.bar[2 * bkt] = 0u;
  if (W > 1 && threadIdx.x < W) st_release_sys(a.signal[threadIdx.x] + flag_pub_idx(bkt, rank), epoch);
  for (int k = a.key_begin + (int)threadIdx.x; k < a.key_end; k += blockDim.x) a.norms_sq[k] = 0.f;
  if (a.last && threadIdx.x == 0) {
    *a.step = *(const volatile int32_t*)a.step + 1;
    __threadfence();
    *(volatile uint32_t*)a.epoch = epoch;
  }
  if (tr && threadIdx.x == 0) tr[5] = globaltimer_ns();
}

__global__ void stamp_kernel(unsigned long long* dst) { *dst = globaltimer_ns(); }
cudaError_t launch_stamp(unsigned long long* dst, cudaStream_t s) {
  stamp_kernel<<<1, 1, 0, s>>>(dst);
  return cudaGetLastError();
}

cudaError_t launch_fused_bucket(const FusedCommArgs& a, int mode, cudaStream_t s) {
  const int n_items = a.blk_end - a.blk_begin;
  cudaError_t e = cudaSuccess;
  if (mode != 2 && (a.world > 1 || a.trace)) {
    bucket_signal_kernel<<<1, 32, 0, s>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (mode == 1) return e;
  if (a.world > 1 || n_items == 0) {
    bucket_wait_kernel<<<1, 32, 0, s>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess || n_items == 0) return e;   // a rank that owns nothing of the bucket is done
  }
  bucket_reduce_kernel<<<n_items, 256, 0, s>>>(a);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  bucket_update_kernel<<<n_items, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_gate_zero(const GateArgs& g, float* grad, const int64_t* ext_off, const int64_t* ext_len,
                             int ext_begin, int ext_end, int grid, cudaStream_t s) {
  bucket_gate_zero_kernel<<<grid, 256, 0, s>>>(g, grad, ext_off, ext_len, ext_begin, ext_end);
  return cudaGetLastError();
}

// Stand-alone gate for consumers that have no in-kernel gate (library GEMM fallbacks, host reads).
__global__ void gate_wait_kernel(GateArgs g) { gate_wait_warp(g); }
cudaError_t launch_gate_wait(const GateArgs& g, cudaStream_t s) {
  gate_wait_kernel<<<1, 32, 0, s>>>(g);
  return cudaGetLastError();
}

// Device-scope barrier on a monotonic counter (stand-alone collectives below).
__device__ __forceinline__ bool grid_barrier_counter(uint32_t* counter, uint32_t target, uint64_t timeout_ns) {
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
  if (!grid_barrier_counter(a.bar_counter, bar_base + gridDim.x, a.timeout_ns)) { if (threadIdx.x == 0) atomicExch(a.error, 2); return; }
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
  if (!grid_barrier_counter(a.bar_counter, bar_base + gridDim.x, a.timeout_ns)) { if (threadIdx.x == 0) atomicExch(a.error, 2); return; }
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
