// torch.ops.srb.tc_gemm: generic entry to the tcgen05 GEMM family (gemm_tcgen05.cu).
// The per-layer wrappers (window maxout fwd, window dX, dW ...) live in Python
// (ops/b200_ops.py) and only fill in the shift tables / epilogue selectors.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include "gemm.h"
#include <cstdlib>

#include "gemm_launch.h"

namespace srb {
namespace {

using at::Tensor;

int num_sms_for(int device) {
  static int cache[64] = {0};
  if (device < 0 || device >= 64) device = 0;
  if (cache[device] == 0) cudaDeviceGetAttribute(&cache[device], cudaDevAttrMultiProcessorCount, device);
  return cache[device];
}

void tc_gemm(const Tensor& A, const Tensor& B, Tensor out, int64_t mode, int64_t epi, int64_t block_n, int64_t M,
             int64_t N, int64_t K, std::vector<int64_t> a_row_shift, std::vector<int64_t> a_col_off,
             std::vector<int64_t> b_row_off, std::vector<int64_t> b_col_off, int64_t splits, int64_t win_w,
             const c10::optional<Tensor>& bias, const c10::optional<Tensor>& which,
             const c10::optional<Tensor>& add_src, const c10::optional<Tensor>& row_scale,
             const c10::optional<Tensor>& m_dev, int64_t max_ctas, int64_t cluster, std::vector<int64_t> gate) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && out.is_cuda());
  TORCH_CHECK(A.scalar_type() == at::kBFloat16 && B.scalar_type() == at::kBFloat16, "tc_gemm: bf16 operands");
  TORCH_CHECK(A.dim() == 2 && B.dim() == 2 && A.stride(1) == 1 && B.stride(1) == 1, "tc_gemm: row-major 2D operands");
  TORCH_CHECK((A.stride(0) * 2) % 16 == 0 && (B.stride(0) * 2) % 16 == 0, "tc_gemm: row pitch must be 16B aligned");
  TORCH_CHECK(((uintptr_t)A.data_ptr() % 16) == 0 && ((uintptr_t)B.data_ptr() % 16) == 0, "tc_gemm: 16B-aligned bases");
  TORCH_CHECK(N % 16 == 0 && (epi != EPI_MAXOUT3 || N % block_n == 0),
              "tc_gemm: N must be a multiple of 16 (of block_n for the maxout epilogue)");
  TORCH_CHECK(N <= 4096 || !(bias.has_value() && bias->defined()), "tc_gemm: bias supported for N <= 4096");
  const int n_shifts = (int)a_row_shift.size();
  TORCH_CHECK(n_shifts >= 1 && n_shifts <= 3 && a_col_off.size() == a_row_shift.size() &&
              b_row_off.size() == a_row_shift.size() && b_col_off.size() == a_row_shift.size());
  c10::cuda::CUDAGuard guard(A.device());
  if ((cluster != 2 && cluster != 3) || !gemm_supports_cluster((int)block_n, (int)mode, (int)epi)) cluster = 1;
  // window GEMM (shifts -1/0/+1 of the same A columns): load A once per k-block with a one-row halo
  bool halo = false;
  if (n_shifts == 3 && splits <= 1 && gemm_supports_halo((int)block_n, (int)mode, (int)epi, (int)cluster) &&
      !(std::getenv("SRB_GEMM_HALO") && std::getenv("SRB_GEMM_HALO")[0] == '0')) {
    bool seen[3] = {false, false, false};
    halo = a_col_off[0] == a_col_off[1] && a_col_off[1] == a_col_off[2];
    for (int s = 0; s < 3 && halo; ++s) {
      const int64_t r = a_row_shift[s] + 1;
      if (r < 0 || r > 2 || seen[r]) halo = false; else seen[r] = true;
    }
  }
  const uint32_t a_rows = halo ? (uint32_t)kHaloRows : 128u;
  CUtensorMap ta, tb;
  int r1, r2;
  if (mode == MODE_KK) {
    r1 = make_tmap_2d_bf16(&ta, A.data_ptr(), (uint64_t)A.size(1), (uint64_t)A.size(0), (uint64_t)A.stride(0) * 2, 64, a_rows);
    // with a 2-CTA cluster each CTA loads (and multicasts) half of the B rows of a tile
    r2 = make_tmap_2d_bf16(&tb, B.data_ptr(), (uint64_t)B.size(1), (uint64_t)B.size(0), (uint64_t)B.stride(0) * 2, 64,
                           (uint32_t)(cluster > 1 ? block_n / 2 : block_n));
  } else if (mode == MODE_KMN) {
    TORCH_CHECK(block_n % 64 == 0, "tc_gemm: MN-major B needs block_n % 64 == 0");
    r1 = make_tmap_2d_bf16(&ta, A.data_ptr(), (uint64_t)A.size(1), (uint64_t)A.size(0), (uint64_t)A.stride(0) * 2, 64, a_rows);
    r2 = make_tmap_2d_bf16(&tb, B.data_ptr(), (uint64_t)B.size(1), (uint64_t)B.size(0), (uint64_t)B.stride(0) * 2, 64, 64);
  } else {
    TORCH_CHECK(block_n % 64 == 0, "tc_gemm: MN-major mode needs block_n % 64 == 0");
    r1 = make_tmap_2d_bf16(&ta, A.data_ptr(), (uint64_t)A.size(1), (uint64_t)A.size(0), (uint64_t)A.stride(0) * 2, 64, 64);
    r2 = make_tmap_2d_bf16(&tb, B.data_ptr(), (uint64_t)B.size(1), (uint64_t)B.size(0), (uint64_t)B.stride(0) * 2, 64, 64);
  }
  TORCH_CHECK(r1 == 0 && r2 == 0, "tc_gemm: cuTensorMapEncodeTiled failed (", r1, ", ", r2, ")");
  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.n_shifts = n_shifts;
  for (int s = 0; s < n_shifts; ++s) {
    p.a_row_shift[s] = (int)a_row_shift[s]; p.a_col_off[s] = (int)a_col_off[s];
    p.b_row_off[s] = (int)b_row_off[s]; p.b_col_off[s] = (int)b_col_off[s];
  }
  p.splits = (int)splits; p.win_w = (int)win_w;
  p.halo = halo ? 1 : 0;
  p.m_dev = m_dev.has_value() && m_dev->defined() ? m_dev->data_ptr<int>() : nullptr;
  p.out = out.data_ptr();
  p.ldo = (int)out.stride(0);
  if (epi == EPI_ATOMIC_F32) {
    TORCH_CHECK(out.scalar_type() == at::kFloat, "tc_gemm: atomic epilogue needs fp32 out");
  } else {
    TORCH_CHECK(out.scalar_type() == at::kBFloat16, "tc_gemm: bf16 out expected");
  }
  p.bias = bias.has_value() && bias->defined() ? (const __nv_bfloat16*)bias->data_ptr() : nullptr;
  p.which = which.has_value() && which->defined() ? which->data_ptr<uint8_t>() : nullptr;
  if (epi == EPI_MAXOUT3) {
    TORCH_CHECK(p.which != nullptr, "tc_gemm: maxout epilogue needs `which`");
  }
  p.add_src = add_src.has_value() && add_src->defined() ? (const __nv_bfloat16*)add_src->data_ptr() : nullptr;
  p.ld_add = p.add_src ? (int)add_src->stride(0) : 0;
  p.row_scale = row_scale.has_value() && row_scale->defined() ? row_scale->data_ptr<float>() : nullptr;
  p.gate = make_gate_args(gate.data(), gate.size());
  int sms = num_sms_for(A.get_device());
  if (max_ctas > 0 && max_ctas < sms) sms = (int)max_ctas;
  cudaError_t e = launch_gemm(ta, tb, p, (int)block_n, (int)mode, (int)epi, (int)cluster, sms,
                              at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(e == cudaSuccess, "tc_gemm launch failed: ", cudaGetErrorString(e), " (block_n=", block_n,
              " mode=", mode, " epi=", epi, ")");
}

}  // namespace

void register_gemm_ops(torch::Library& m) {
  m.def(
      "tc_gemm(Tensor A, Tensor B, Tensor(a!) out, int mode, int epi, int block_n, int M, int N, int K, "
      "int[] a_row_shift, int[] a_col_off, int[] b_row_off, int[] b_col_off, int splits, int win_w, "
      "Tensor? bias, Tensor? which, Tensor? add_src, Tensor? row_scale, Tensor? m_dev, int max_ctas, int cluster, "
      "int[] gate) -> ()");
}
void register_gemm_impls(torch::Library& m) { m.impl("tc_gemm", tc_gemm); }

}  // namespace srb
