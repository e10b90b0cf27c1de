// torch.ops.srb.tc_gemm: generic entry to the tcgen05 GEMM family (gemm_tcgen05.cu).
// The per-layer wrappers (window maxout fwd, window dX, dW ...) live in Python
// (ops/b200_ops.py) and only fill in the shift tables / epilogue selectors.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include "gemm.h"
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

void tc_gemm_impl(const Tensor& A, const Tensor& B, Tensor out, int64_t mode, int64_t epi, int64_t block_n, int64_t M,
                  int64_t N, int64_t K, std::vector<int64_t> a_row_shift, std::vector<int64_t> a_col_off,
                  std::vector<int64_t> b_row_off, std::vector<int64_t> b_col_off, int64_t splits, int64_t win_w,
                  const c10::optional<Tensor>& bias, const c10::optional<Tensor>& which,
                  const c10::optional<Tensor>& add_src, const c10::optional<Tensor>& row_scale,
                  const c10::optional<Tensor>& m_dev, int64_t max_ctas, int64_t cluster, const LnFuse* ln) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && out.is_cuda());
  TORCH_CHECK(A.scalar_type() == at::kBFloat16 && B.scalar_type() == at::kBFloat16, "tc_gemm: bf16 operands");
  TORCH_CHECK(A.dim() == 2 && B.dim() == 2 && A.stride(1) == 1 && B.stride(1) == 1, "tc_gemm: row-major 2D operands");
  TORCH_CHECK((A.stride(0) * 2) % 16 == 0 && (B.stride(0) * 2) % 16 == 0, "tc_gemm: row pitch must be 16B aligned");
  TORCH_CHECK(((uintptr_t)A.data_ptr() % 16) == 0 && ((uintptr_t)B.data_ptr() % 16) == 0, "tc_gemm: 16B-aligned bases");
  TORCH_CHECK(N % block_n == 0, "tc_gemm: N must be a multiple of block_n");
  TORCH_CHECK(N <= 4096 || !(bias.has_value() && bias->defined()), "tc_gemm: bias supported for N <= 4096");
  const int n_shifts = (int)a_row_shift.size();
  TORCH_CHECK(n_shifts >= 1 && n_shifts <= 3 && a_col_off.size() == a_row_shift.size() &&
              b_row_off.size() == a_row_shift.size() && b_col_off.size() == a_row_shift.size());
  c10::cuda::CUDAGuard guard(A.device());
  if ((cluster != 2 && cluster != 3) || !gemm_supports_cluster((int)block_n, (int)mode, (int)epi)) cluster = 1;
  CUtensorMap ta, tb;
  int r1, r2;
  if (mode == MODE_KK) {
    r1 = make_tmap_2d_bf16(&ta, A.data_ptr(), (uint64_t)A.size(1), (uint64_t)A.size(0), (uint64_t)A.stride(0) * 2, 64, 128);
    // with a 2-CTA cluster each CTA loads (and multicasts) half of the B rows of a tile
    r2 = make_tmap_2d_bf16(&tb, B.data_ptr(), (uint64_t)B.size(1), (uint64_t)B.size(0), (uint64_t)B.stride(0) * 2, 64,
                           (uint32_t)(cluster > 1 ? block_n / 2 : block_n));
  } else if (mode == MODE_KMN) {
    TORCH_CHECK(block_n % 64 == 0, "tc_gemm: MN-major B needs block_n % 64 == 0");
    r1 = make_tmap_2d_bf16(&ta, A.data_ptr(), (uint64_t)A.size(1), (uint64_t)A.size(0), (uint64_t)A.stride(0) * 2, 64, 128);
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
  if (epi == EPI_MAXOUT3 || epi == EPI_MAXOUT3_LN) {
    TORCH_CHECK(p.which != nullptr, "tc_gemm: maxout epilogue needs `which`");
  }
  if (epi == EPI_MAXOUT3_LN) {
    TORCH_CHECK(ln != nullptr && (128 % (N / block_n)) == 0 && (128 / (N / block_n)) % 4 == 0 && N / 3 == 256,
                "tc_gemm: fused LayerNorm needs width 256 and an N tile count dividing 32");
    p.ln = *ln;
  }
  p.add_src = add_src.has_value() && add_src->defined() ? (const __nv_bfloat16*)add_src->data_ptr() : nullptr;
  p.ld_add = p.add_src ? (int)add_src->stride(0) : 0;
  p.row_scale = row_scale.has_value() && row_scale->defined() ? row_scale->data_ptr<float>() : nullptr;
  int sms = num_sms_for(A.get_device());
  if (max_ctas > 0 && max_ctas < sms) sms = (int)max_ctas;
  cudaError_t e = launch_gemm(ta, tb, p, (int)block_n, (int)mode, (int)epi, (int)cluster, sms,
                              at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(e == cudaSuccess, "tc_gemm launch failed: ", cudaGetErrorString(e), " (block_n=", block_n,
              " mode=", mode, " epi=", epi, ")");
}

void tc_gemm(const Tensor& A, const Tensor& B, Tensor out, int64_t mode, int64_t epi, int64_t block_n, int64_t M,
             int64_t N, int64_t K, std::vector<int64_t> a_row_shift, std::vector<int64_t> a_col_off,
             std::vector<int64_t> b_row_off, std::vector<int64_t> b_col_off, int64_t splits, int64_t win_w,
             const c10::optional<Tensor>& bias, const c10::optional<Tensor>& which,
             const c10::optional<Tensor>& add_src, const c10::optional<Tensor>& row_scale,
             const c10::optional<Tensor>& m_dev, int64_t max_ctas, int64_t cluster) {
  TORCH_CHECK(epi != EPI_MAXOUT3_LN, "tc_gemm: use tc_gemm_maxout_ln for the fused LayerNorm epilogue");
  tc_gemm_impl(A, B, out, mode, epi, block_n, M, N, K, a_row_shift, a_col_off, b_row_off, b_col_off, splits, win_w,
               bias, which, add_src, row_scale, m_dev, max_ctas, cluster, nullptr);
}

// (window) GEMM + bias + maxout(3) -> H, `which`; then, inside the same kernel, LayerNorm / dropout /
// residual / mask of finished rows -> Y, xhat, rstd.  `counters`: int32 (2, >= ceil(M/128)+1), zeroed once.
void tc_gemm_maxout_ln(const Tensor& A, const Tensor& B, Tensor H, Tensor which, const c10::optional<Tensor>& bias,
                       const c10::optional<Tensor>& G, const c10::optional<Tensor>& beta,
                       const c10::optional<Tensor>& Xres, const Tensor& mask, Tensor Y, Tensor xhat, Tensor rstd,
                       Tensor counters, int64_t M, int64_t N, int64_t K, std::vector<int64_t> a_row_shift,
                       std::vector<int64_t> a_col_off, std::vector<int64_t> b_row_off, std::vector<int64_t> b_col_off,
                       double drop_p, int64_t seed, const c10::optional<Tensor>& seed_dev, int64_t cluster) {
  const int64_t nO = N / 3;
  TORCH_CHECK(Y.is_cuda() && Y.is_contiguous() && Y.scalar_type() == at::kBFloat16 && Y.size(1) == nO);
  TORCH_CHECK(xhat.is_contiguous() && xhat.scalar_type() == at::kBFloat16 && rstd.scalar_type() == at::kFloat);
  TORCH_CHECK(H.is_contiguous() && H.size(1) == nO && mask.scalar_type() == at::kFloat && mask.is_contiguous());
  TORCH_CHECK(counters.scalar_type() == at::kInt && counters.dim() == 2 && counters.size(0) == 2 &&
              counters.size(1) >= (M + 127) / 128 + 1 && counters.is_contiguous());
  LnFuse ln{};
  const bool has_ln = G.has_value() && G->defined();
  ln.g = has_ln ? (const __nv_bfloat16*)G->data_ptr() : nullptr;
  ln.beta = has_ln ? (const __nv_bfloat16*)beta->data_ptr() : nullptr;
  if (Xres.has_value() && Xres->defined()) {
    TORCH_CHECK(Xres->is_contiguous() && Xres->size(1) == nO && Xres->scalar_type() == at::kBFloat16);
    ln.xres = (const __nv_bfloat16*)Xres->data_ptr();
  }
  ln.mask = mask.data_ptr<float>();
  ln.y = (__nv_bfloat16*)Y.data_ptr();
  ln.xhat = (__nv_bfloat16*)xhat.data_ptr();
  ln.rstd = rstd.data_ptr<float>();
  ln.done = counters.data_ptr<int>();
  ln.consumed = counters.data_ptr<int>() + counters.size(1);
  ln.drop_p = (float)drop_p;
  ln.seed = (unsigned long long)seed;
  ln.seed_dev = seed_dev.has_value() && seed_dev->defined() ? (const long long*)seed_dev->data_ptr<int64_t>() : nullptr;
  tc_gemm_impl(A, B, H, MODE_KK, EPI_MAXOUT3_LN, 192, M, N, K, a_row_shift, a_col_off, b_row_off, b_col_off, 1, 0, bias,
               which, c10::nullopt, c10::nullopt, c10::nullopt, 0, cluster, &ln);
}

}  // namespace

void register_gemm_ops(torch::Library& m) {
  m.def(
      "tc_gemm_maxout_ln(Tensor A, Tensor B, Tensor(a!) H, Tensor(b!) which, Tensor? bias, Tensor? G, Tensor? beta, "
      "Tensor? Xres, Tensor mask, Tensor(c!) Y, Tensor(d!) xhat, Tensor(e!) rstd, Tensor(f!) counters, int M, int N, "
      "int K, int[] a_row_shift, int[] a_col_off, int[] b_row_off, int[] b_col_off, float drop_p, int seed, "
      "Tensor? seed_dev, int cluster) -> ()");
  m.def(
      "tc_gemm(Tensor A, Tensor B, Tensor(a!) out, int mode, int epi, int block_n, int M, int N, int K, "
      "int[] a_row_shift, int[] a_col_off, int[] b_row_off, int[] b_col_off, int splits, int win_w, "
      "Tensor? bias, Tensor? which, Tensor? add_src, Tensor? row_scale, Tensor? m_dev, int max_ctas, int cluster) -> ()");
}
void register_gemm_impls(torch::Library& m) {
  m.impl("tc_gemm", tc_gemm);
  m.impl("tc_gemm_maxout_ln", tc_gemm_maxout_ln);
}

}  // namespace srb
