// Argument blocks for the peer-memory collective kernels (comm_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srb {

constexpr int kMaxWorld = 16;
constexpr int kCommChunk = 4096;          // elements per work item
enum { kSlotGrad = 0, kSlotRead = 1, kSlotParam = 2, kNumSlots = 3 };

struct FusedCommArgs {
  float* grad[kMaxWorld];        // every rank's flat fp32 gradient buffer (peer-mapped)
  void* param[kMaxWorld];        // every rank's flat bf16 weight buffer (peer-mapped)
  uint32_t* signal[kMaxWorld];   // every rank's signal pad
  const float* grad_mc;          // multicast address of the gradient buffers (or null)
  void* param_mc;                // multicast address of the weight buffers (or null)
  float* master;                 // fp32 master weights of my shard (shard_cap)
  float* m1;
  float* m2;
  float* norms_sq;               // (n_keys) scratch, zero on entry and on exit
  const int32_t* blk_key;        // work item -> key
  const int32_t* blk_off;        // work item -> chunk index inside the key
  const int64_t* key_off;        // key -> offset relative to my shard start
  const int64_t* key_len;        // key -> padded length (multiple of 128)
  const float* hyper;            // {lr, beta1, beta2, eps, grad_clip, l2, l2_is_wd, grad_scale}
  int32_t* step;                 // Adam update counter (device)
  uint32_t* epoch;               // flag epoch (device)
  uint32_t* bar_counter;         // grid barrier counter (device, monotonic)
  int32_t* error;                // 0 = ok, else the phase that timed out
  int64_t shard_start;           // elements
  int64_t shard_cap;
  int64_t total_elems;
  uint64_t timeout_ns;
  int n_blocks, n_keys, world, rank, wait_params;
};

struct P2PCollArgs {
  void* buf[kMaxWorld];          // every rank's symmetric buffer
  uint32_t* signal[kMaxWorld];
  void* mc;                      // multicast address or null
  void* out;                     // local shard (reduce-scatter output / all-gather input)
  uint32_t* epoch;
  uint32_t* bar_counter;
  int32_t* error;
  int64_t shard_elems;           // fp32 elements per rank
  uint64_t timeout_ns;
  int world, rank;
};

cudaError_t launch_fused_rs_adam_ag(const FusedCommArgs& a, int grid, cudaStream_t s);
cudaError_t launch_p2p_reduce_scatter(const P2PCollArgs& a, int grid, cudaStream_t s);
cudaError_t launch_p2p_all_gather(const P2PCollArgs& a, int grid, cudaStream_t s);

}  // namespace srb
