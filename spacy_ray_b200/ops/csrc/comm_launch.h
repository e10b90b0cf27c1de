// Argument blocks for the peer-memory collective kernels (comm_kernels.cu) and the
// consumer-side "published" gate that the tcgen05 GEMM / HashEmbed kernels embed (gate.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace srb {

constexpr int kMaxWorld = 16;
constexpr int kMaxBuckets = 32;
constexpr int kCommChunk = 4096;          // elements per work item
// Symmetric signal page (uint32 words, one page per rank, every rank can write every page):
//   grad_ready[b][r]  word (b * kMaxWorld + r)                   rank r's gradients of bucket b, epoch e, are complete
//   published[b][r]   word ((kMaxBuckets + b) * kMaxWorld + r)   rank r has stored its bucket-b weights of epoch e
//                                                                into THIS rank's weight buffer (and zeroed its shard
//                                                                of this rank's gradient buffer)
// Values are epochs (monotonic), never reset.
constexpr int kSignalWords = 2 * kMaxBuckets * kMaxWorld;
__host__ __device__ __forceinline__ int flag_grad_idx(int b, int r) { return b * kMaxWorld + r; }
__host__ __device__ __forceinline__ int flag_pub_idx(int b, int r) { return (kMaxBuckets + b) * kMaxWorld + r; }

// legacy slots of the stand-alone collectives (their callers pass a private signal page)
enum { kSlotGrad = 0, kSlotRead = 1, kSlotParam = 2, kNumSlots = 3 };
enum { kOptAdam = 0, kOptRAdam = 1, kOptSGD = 2 };

// Consumer-side gate: "every owner's weights of these buckets, epoch *epoch, have landed here".
struct GateArgs {
  const uint32_t* flags;      // my signal page (null = no gate)
  const uint32_t* epoch;      // device word: number of completed exchanges
  int32_t* error;             // timeout -> error code 7
  uint64_t timeout_ns;
  uint32_t mask;              // buckets to wait for
  int world;
};

// host: gate descriptor as passed through the op bindings:
// [flags_ptr, epoch_ptr, error_ptr, bucket_mask, world, timeout_ms]; empty / mask 0 = no gate
inline GateArgs make_gate_args(const int64_t* v, size_t n) {
  GateArgs g{};
  if (n >= 6 && v[0] != 0 && v[3] != 0) {
    g.flags = reinterpret_cast<const uint32_t*>(v[0]);
    g.epoch = reinterpret_cast<const uint32_t*>(v[1]);
    g.error = reinterpret_cast<int32_t*>(v[2]);
    g.mask = (uint32_t)v[3];
    g.world = (int)v[4];
    g.timeout_ns = (uint64_t)v[5] * 1000000ull;
  }
  return g;
}

struct FusedCommArgs {
  float* grad[kMaxWorld];        // every rank's flat fp32 gradient buffer (peer-mapped)
  void* param[kMaxWorld];        // every rank's flat bf16 weight buffer (peer-mapped)
  uint32_t* signal[kMaxWorld];   // every rank's signal page
  float* grad_mc;                // multicast address of the gradient buffers (or null)
  void* param_mc;                // multicast address of the weight buffers (or null)
  float* red;                    // (shard_cap) reduced gradient of my shard (local scratch; == my grad shard if world == 1)
  float* master;                 // fp32 master weights of my shard (shard_cap)
  float* m1;
  float* m2;
  float* avg;                    // parameter averages of my shard (or null)
  float* norms_sq;               // (n_keys) scratch, zero on entry and on exit
  const int32_t* blk_key;        // work item -> key
  const int32_t* blk_off;        // work item -> chunk index inside the key
  const int64_t* key_off;        // key -> offset relative to my shard start
  const int64_t* key_len;        // key -> padded length (multiple of 128)
  const float* hyper;            // {lr, beta1, beta2, eps, grad_clip, l2, l2_is_wd, grad_scale}
  int32_t* step;                 // optimizer update counter (device)
  uint32_t* epoch;               // exchange epoch (device)
  uint32_t* bar;                 // grid barrier state: 2 words (count, generation) per bucket
  int32_t* error;                // 0 = ok, else the phase that timed out
  int64_t shard_start;           // elements
  uint64_t timeout_ns;
  int blk_begin, blk_end;        // this bucket's work items
  int key_begin, key_end;        // this bucket's owned keys
  int bucket;
  int last;                      // 1 = this launch ends the step (advances epoch / update counter)
  int world, rank;
  int opt_mode;                  // kOptAdam / kOptRAdam / kOptSGD
  uint64_t test_delay_ns;        // tests only: hold the weight publication back by this long (late publisher)
  // optional %globaltimer trace (null = off), 8 words per bucket:
  //   [0] signal kernel start  [1] all peers' gradients seen  [2] first reduce CTA start  [3] last reduce CTA end
  //   [4] first update CTA start  [5] bucket published (last update CTA)  [6..7] spare
  unsigned long long* trace;
};
constexpr int kTraceWords = 8;

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

// mode 0: the whole bucket pipeline on one stream; 1: only the grad-ready signal (own stream, so a
// bucket's signal never queues behind an earlier bucket's exchange); 2: wait for the peers + reduce + update
cudaError_t launch_fused_bucket(const FusedCommArgs& a, int mode, cudaStream_t s);
// wait until every owner has published bucket(s) `g.mask` (= they are done reading my gradients of
// that bucket), then clear my gradient accumulators of the bucket: extents [ext_begin, ext_end)
cudaError_t launch_gate_zero(const GateArgs& g, float* grad, const int64_t* ext_off, const int64_t* ext_len,
                             int ext_begin, int ext_end, int grid, cudaStream_t s);
cudaError_t launch_gate_wait(const GateArgs& g, cudaStream_t s);
cudaError_t launch_stamp(unsigned long long* dst, cudaStream_t s);     // *dst = %globaltimer (one thread)
cudaError_t launch_p2p_reduce_scatter(const P2PCollArgs& a, int grid, cudaStream_t s);
cudaError_t launch_p2p_all_gather(const P2PCollArgs& a, int grid, cudaStream_t s);

}  // namespace srb
