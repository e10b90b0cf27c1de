// torch.ops.srb.{fused_comm_bucket, gate_wait, p2p_collective}
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include "comm.h"
#include "comm_launch.h"

namespace srb {
namespace {
using at::Tensor;

// One bucket of the gradient exchange (see comm_kernels.cu).  Launched on the CURRENT stream.
void fused_comm_bucket(std::vector<int64_t> grad_ptrs, std::vector<int64_t> param_ptrs, std::vector<int64_t> signal_ptrs,
                       int64_t grad_mc, int64_t param_mc, Tensor red, Tensor master, Tensor m1, Tensor m2,
                       const c10::optional<Tensor>& avg, Tensor norms, const Tensor& blk_key, const Tensor& blk_off,
                       const Tensor& key_off, const Tensor& key_len, const Tensor& hyper, Tensor step, Tensor epoch,
                       Tensor bar, Tensor error, int64_t shard_start, int64_t blk_begin, int64_t blk_end,
                       int64_t key_begin, int64_t key_end, int64_t bucket, bool last, int64_t rank, int64_t grid,
                       int64_t opt_mode, double timeout_s, int64_t test_delay_us, const c10::optional<Tensor>& trace) {
  const int W = (int)grad_ptrs.size();
  TORCH_CHECK(W >= 1 && W <= kMaxWorld && (int)param_ptrs.size() == W && (int)signal_ptrs.size() == W);
  TORCH_CHECK(master.is_cuda() && master.scalar_type() == at::kFloat && red.scalar_type() == at::kFloat);
  TORCH_CHECK(shard_start % 4 == 0 && bucket >= 0 && bucket < kMaxBuckets && grid >= 1);
  TORCH_CHECK(bar.numel() >= 2 * kMaxBuckets, "fused_comm_bucket: barrier state needs 2 words per bucket");
  TORCH_CHECK(blk_begin >= 0 && blk_end <= blk_key.numel() && key_end <= key_off.numel());
  c10::cuda::CUDAGuard guard(master.device());
  FusedCommArgs a{};
  for (int p = 0; p < W; ++p) {
    a.grad[p] = reinterpret_cast<float*>(grad_ptrs[p]);
    a.param[p] = reinterpret_cast<void*>(param_ptrs[p]);
    a.signal[p] = reinterpret_cast<uint32_t*>(signal_ptrs[p]);
  }
  a.grad_mc = reinterpret_cast<float*>(grad_mc);
  a.param_mc = reinterpret_cast<void*>(param_mc);
  a.red = red.data_ptr<float>();
  a.master = master.data_ptr<float>(); a.m1 = m1.data_ptr<float>(); a.m2 = m2.data_ptr<float>();
  a.avg = avg.has_value() && avg->defined() ? avg->data_ptr<float>() : nullptr;
  a.norms_sq = norms.data_ptr<float>();
  a.blk_key = blk_key.data_ptr<int32_t>(); a.blk_off = blk_off.data_ptr<int32_t>();
  a.key_off = key_off.data_ptr<int64_t>(); a.key_len = key_len.data_ptr<int64_t>();
  a.hyper = hyper.data_ptr<float>();
  a.step = step.data_ptr<int32_t>();
  a.epoch = reinterpret_cast<uint32_t*>(epoch.data_ptr<int32_t>());
  a.bar = reinterpret_cast<uint32_t*>(bar.data_ptr<int32_t>());
  a.error = error.data_ptr<int32_t>();
  a.shard_start = shard_start;
  a.timeout_ns = (uint64_t)(timeout_s * 1e9);
  a.blk_begin = (int)blk_begin; a.blk_end = (int)blk_end; a.key_begin = (int)key_begin; a.key_end = (int)key_end;
  a.bucket = (int)bucket; a.last = last ? 1 : 0;
  a.world = W; a.rank = (int)rank; a.opt_mode = (int)opt_mode;
  a.test_delay_ns = (uint64_t)test_delay_us * 1000ull;
  if (trace.has_value() && trace->defined()) {
    TORCH_CHECK(trace->scalar_type() == at::kLong && trace->numel() >= kMaxBuckets * kTraceWords);
    a.trace = reinterpret_cast<unsigned long long*>(trace->data_ptr<int64_t>());
  }
  // `grid` carries the launch mode: 0 = whole pipeline, 1 = grad-ready signal only, 2 = wait + reduce + update
  cudaError_t e = launch_fused_bucket(a, (int)grid, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(e == cudaSuccess, "fused_comm_bucket launch failed: ", cudaGetErrorString(e));
}

// Wait for the owners' "published" flags of the gate's buckets, then clear my gradient accumulators
// of those buckets (see bucket_gate_zero_kernel).
void gate_zero(Tensor grad, std::vector<int64_t> gate, const Tensor& ext_off, const Tensor& ext_len, int64_t ext_begin,
               int64_t ext_end, int64_t grid) {
  TORCH_CHECK(grad.is_cuda() && grad.scalar_type() == at::kFloat && ext_off.scalar_type() == at::kLong &&
              ext_len.scalar_type() == at::kLong && ext_end <= ext_off.numel() && grid >= 1);
  c10::cuda::CUDAGuard guard(grad.device());
  GateArgs g = make_gate_args(gate.data(), gate.size());
  cudaError_t e = launch_gate_zero(g, grad.data_ptr<float>(), ext_off.data_ptr<int64_t>(), ext_len.data_ptr<int64_t>(),
                                   (int)ext_begin, (int)ext_end, (int)grid, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(e == cudaSuccess, "gate_zero launch failed: ", cudaGetErrorString(e));
}

// trace[index] = %globaltimer on the current stream (one-thread kernel): step / phase boundaries for the
// overlap trace (benchmarks/exchange_trace.py)
void stamp(Tensor trace, int64_t index) {
  TORCH_CHECK(trace.is_cuda() && trace.scalar_type() == at::kLong && index >= 0 && index < trace.numel());
  c10::cuda::CUDAGuard guard(trace.device());
  cudaError_t e = launch_stamp(reinterpret_cast<unsigned long long*>(trace.data_ptr<int64_t>()) + index,
                               at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(e == cudaSuccess, "stamp launch failed: ", cudaGetErrorString(e));
}

// Stand-alone consumer gate (one warp): used in front of consumers without an in-kernel gate.
void gate_wait(const Tensor& epoch, std::vector<int64_t> gate) {
  GateArgs g = make_gate_args(gate.data(), gate.size());
  if (!g.flags) return;
  c10::cuda::CUDAGuard guard(epoch.device());
  cudaError_t e = launch_gate_wait(g, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(e == cudaSuccess, "gate_wait launch failed: ", cudaGetErrorString(e));
}

void p2p_collective(int64_t kind, std::vector<int64_t> buf_ptrs, std::vector<int64_t> signal_ptrs, int64_t mc,
                    Tensor local, Tensor epoch, Tensor bar_counter, Tensor error, int64_t shard_elems, int64_t rank,
                    int64_t grid, double timeout_s) {
  const int W = (int)buf_ptrs.size();
  TORCH_CHECK(W >= 1 && W <= kMaxWorld && shard_elems % 4 == 0);
  c10::cuda::CUDAGuard guard(local.device());
  P2PCollArgs a{};
  for (int p = 0; p < W; ++p) {
    a.buf[p] = reinterpret_cast<void*>(buf_ptrs[p]);
    a.signal[p] = reinterpret_cast<uint32_t*>(signal_ptrs[p]);
  }
  a.mc = reinterpret_cast<void*>(mc);
  a.out = local.data_ptr();
  a.epoch = reinterpret_cast<uint32_t*>(epoch.data_ptr<int32_t>());
  a.bar_counter = reinterpret_cast<uint32_t*>(bar_counter.data_ptr<int32_t>());
  a.error = error.data_ptr<int32_t>();
  a.shard_elems = shard_elems; a.timeout_ns = (uint64_t)(timeout_s * 1e9);
  a.world = W; a.rank = (int)rank;
  auto s = at::cuda::getCurrentCUDAStream().stream();
  cudaError_t e = kind == 0 ? launch_p2p_reduce_scatter(a, (int)grid, s) : launch_p2p_all_gather(a, (int)grid, s);
  TORCH_CHECK(e == cudaSuccess, "p2p collective launch failed: ", cudaGetErrorString(e));
}

}  // namespace

void register_comm_ops(torch::Library& m) {
  m.def(
      "fused_comm_bucket(int[] grad_ptrs, int[] param_ptrs, int[] signal_ptrs, int grad_mc, int param_mc, "
      "Tensor(r!) red, Tensor(a!) master, Tensor(b!) m1, Tensor(c!) m2, Tensor(i!)? avg, Tensor(d!) norms, "
      "Tensor blk_key, Tensor blk_off, Tensor key_off, Tensor key_len, Tensor hyper, Tensor(e!) step, "
      "Tensor(f!) epoch, Tensor(g!) bar, Tensor(h!) error, int shard_start, int blk_begin, int blk_end, "
      "int key_begin, int key_end, int bucket, bool last, int rank, int grid, int opt_mode, float timeout_s, "
      "int test_delay_us, Tensor(t!)? trace) -> ()");
  m.def("stamp(Tensor(a!) trace, int index) -> ()");
  m.def("gate_zero(Tensor(a!) grad, int[] gate, Tensor ext_off, Tensor ext_len, int ext_begin, int ext_end, int grid) -> ()");
  m.def("gate_wait(Tensor epoch, int[] gate) -> ()");
  m.def(
      "p2p_collective(int kind, int[] buf_ptrs, int[] signal_ptrs, int mc, Tensor(a!) local, Tensor(b!) epoch, "
      "Tensor(c!) bar_counter, Tensor(d!) error, int shard_elems, int rank, int grid, float timeout_s) -> ()");
}
void register_comm_impls(torch::Library& m) {
  m.impl("fused_comm_bucket", fused_comm_bucket);
  m.impl("gate_wait", gate_wait);
  m.impl("stamp", stamp);
  m.impl("gate_zero", gate_zero);
  m.impl("p2p_collective", p2p_collective);
}

}  // namespace srb
