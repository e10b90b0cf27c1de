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

void tc_gemm_impl(const Tensor& A, const Tensor& B, Tensor out, int64_t mode, int64_t epi, int64_t block_n, int64_t M,
                  int64_t N, int64_t K, std::vector<int64_t> a_row_shift, std::vector<int64_t> a_col_off,
                  std::vector<int64_t> b_row_off, std::vector<int64_t> b_col_off, int64_t splits, int64_t win_w,
                  const c10::optional<Tensor>& bias, const c10::optional<Tensor>& which,
                  const c10::optional<Tensor>& add_src, const c10::optional<Tensor>& row_scale,
                  const c10::optional<Tensor>& m_dev, int64_t max_ctas, int64_t cluster, std::vector<int64_t> gate,
                  const LnArgs* ln) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && out.is_cuda());
  TORCH_CHECK(A.scalar_type() == at::kBFloat16 && B.scalar_type() == at::kBFloat16, "tc_gemm: bf16 operands");
  TORCH_CHECK(A.dim() == 2 && B.dim() == 2 && A.stride(1) == 1 && B.stride(1) == 1, "tc_gemm: row-major 2D operands");
  TORCH_CHECK((A.stride(0) * 2) % 16 == 0 && (B.stride(0) * 2) % 16 == 0, "tc_gemm: row pitch must be 16B aligned");
  TORCH_CHECK(((uintptr_t)A.data_ptr() % 16) == 0 && ((uintptr_t)B.data_ptr() % 16) == 0, "tc_gemm: 16B-aligned bases");
  TORCH_CHECK(N % 16 == 0 && ((epi != EPI_MAXOUT3 && epi != EPI_MAXOUT3_LN) || N % block_n == 0),
              "tc_gemm: N must be a multiple of 16 (of block_n for the maxout epilogue)");
  TORCH_CHECK(N <= 4096 || !(bias.has_value() && bias->defined()), "tc_gemm: bias supported for N <= 4096");
  const int n_shifts = (int)a_row_shift.size();
  TORCH_CHECK(n_shifts >= 1 && n_shifts <= 3 && a_col_off.size() == a_row_shift.size() &&
              b_row_off.size() == a_row_shift.size() && b_col_off.size() == a_row_shift.size());
  c10::cuda::CUDAGuard guard(A.device());
  if ((cluster != 2 && cluster != 3) || !gemm_supports_cluster((int)block_n, (int)mode, (int)epi)) cluster = 1;
  TORCH_CHECK(epi != EPI_MAXOUT3_LN || (ln != nullptr && cluster == 3 && block_n == 192 && mode == MODE_KK),
              "tc_gemm: the fused LayerNorm epilogue exists for the pair-MMA cluster, 192-column tiles, K-major operands");
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
  if (epi == EPI_MAXOUT3 || epi == EPI_MAXOUT3_LN) {
    TORCH_CHECK(p.which != nullptr, "tc_gemm: maxout epilogue needs `which`");
  }
  if (ln) p.ln = *ln;
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

void tc_gemm(const Tensor& A, const Tensor& B, Tensor out, int64_t mode, int64_t epi, int64_t block_n, int64_t M,
             int64_t N, int64_t K, std::vector<int64_t> a_row_shift, std::vector<int64_t> a_col_off,
             std::vector<int64_t> b_row_off, std::vector<int64_t> b_col_off, int64_t splits, int64_t win_w,
             const c10::optional<Tensor>& bias, const c10::optional<Tensor>& which,
             const c10::optional<Tensor>& add_src, const c10::optional<Tensor>& row_scale,
             const c10::optional<Tensor>& m_dev, int64_t max_ctas, int64_t cluster, std::vector<int64_t> gate) {
  TORCH_CHECK(epi != EPI_MAXOUT3_LN, "tc_gemm: use tc_gemm_maxout_ln for the fused LayerNorm epilogue");
  tc_gemm_impl(A, B, out, mode, epi, block_n, M, N, K, a_row_shift, a_col_off, b_row_off, b_col_off, splits, win_w, bias,
               which, add_src, row_scale, m_dev, max_ctas, cluster, gate, nullptr);
}

// Y = mask * (dropout(LayerNorm(maxout3(A (*) B^T + bias))) + Xres): the (window) GEMM with the whole
// rest of the layer in its epilogue (EPI_MAXOUT3_LN, gemm_launch.h).  `stats` (fp32, >= rows_pad * n_tiles * 4
// with rows_pad = M rounded up to 256, zero-initialised ONCE) and `seq` (int32 [tag = 1, 0, 0]: launch tag, finished-CTA
// count, time-out flag; maintained by the kernel) are caller-owned persistent scratch.
void tc_gemm_maxout_ln(const Tensor& A, const Tensor& B, Tensor Y, Tensor which, Tensor xhat, Tensor rstd,
                       const Tensor& bias, const Tensor& G, const Tensor& beta, const c10::optional<Tensor>& xres,
                       const Tensor& mask, Tensor stats, Tensor cnt, int64_t M, int64_t N, int64_t K,
                       std::vector<int64_t> a_row_shift, std::vector<int64_t> a_col_off, std::vector<int64_t> b_row_off,
                       std::vector<int64_t> b_col_off, double drop_p, int64_t seed, const c10::optional<Tensor>& seed_dev,
                       const c10::optional<Tensor>& m_dev, int64_t cluster, std::vector<int64_t> gate) {
  const int64_t nO = N / 3, n_tiles = N / 192;
  TORCH_CHECK(N % 192 == 0 && n_tiles <= 8 && N + 2 * nO + 512 <= 4096, "tc_gemm_maxout_ln: 3 * nO must be a multiple of 192, nO <= 512");
  TORCH_CHECK(Y.scalar_type() == at::kBFloat16 && xhat.scalar_type() == at::kBFloat16 && Y.size(1) == nO &&
              xhat.is_contiguous() && xhat.size(1) == nO && which.is_contiguous() && which.scalar_type() == at::kByte &&
              which.size(1) == nO && rstd.scalar_type() == at::kFloat && mask.scalar_type() == at::kFloat);
  TORCH_CHECK(G.scalar_type() == at::kBFloat16 && beta.scalar_type() == at::kBFloat16 && G.numel() == nO &&
              beta.numel() == nO && G.is_contiguous() && beta.is_contiguous());
  const int64_t rows_pad = (M + 255) / 256 * 256;
  TORCH_CHECK(stats.scalar_type() == at::kFloat && stats.numel() >= rows_pad * n_tiles * 4 && stats.is_contiguous() &&
              ((uintptr_t)stats.data_ptr() % 16) == 0, "tc_gemm_maxout_ln: stats scratch too small");
  TORCH_CHECK(stats.numel() >= rows_pad * n_tiles * 4, "tc_gemm_maxout_ln: stats scratch too small");
  TORCH_CHECK(cnt.scalar_type() == at::kInt && cnt.numel() >= 3 && cnt.is_contiguous(), "tc_gemm_maxout_ln: seq scratch");
  TORCH_CHECK(Y.size(0) >= M && xhat.size(0) >= M && which.size(0) >= M && rstd.numel() >= M && mask.numel() >= M);
  LnArgs ln{};
  ln.G = (const __nv_bfloat16*)G.data_ptr();
  ln.beta = (const __nv_bfloat16*)beta.data_ptr();
  const bool has_res = xres.has_value() && xres->defined();
  if (has_res) {
    TORCH_CHECK(xres->scalar_type() == at::kBFloat16 && xres->size(1) == nO && xres->stride(1) == 1 &&
                (xres->stride(0) * 2) % 16 == 0 && xres->size(0) >= M);
  }
  ln.xres = has_res ? (const __nv_bfloat16*)xres->data_ptr() : nullptr;
  ln.ld_res = has_res ? (int)xres->stride(0) : 0;
  ln.mask = mask.data_ptr<float>();
  ln.xhat = (__nv_bfloat16*)xhat.data_ptr();
  ln.rstd = rstd.data_ptr<float>();
  ln.stats = (float4*)stats.data_ptr<float>();
  ln.seq = (unsigned int*)cnt.data_ptr<int>();
  ln.drop_p = (float)drop_p;
  ln.seed = (uint64_t)seed;
  ln.seed_dev = seed_dev.has_value() && seed_dev->defined() ? seed_dev->data_ptr<int64_t>() : nullptr;
  tc_gemm_impl(A, B, Y, MODE_KK, EPI_MAXOUT3_LN, 192, M, N, K, a_row_shift, a_col_off, b_row_off, b_col_off, 1, 0, bias,
               which, c10::nullopt, c10::nullopt, m_dev, 0, cluster, gate, &ln);
}

}  // namespace

void register_gemm_ops(torch::Library& m) {
  m.def(
      "tc_gemm_maxout_ln(Tensor A, Tensor B, Tensor(a!) Y, Tensor(b!) which, Tensor(c!) xhat, Tensor(d!) rstd, Tensor bias, "
      "Tensor G, Tensor beta, Tensor? xres, Tensor mask, Tensor(e!) stats, Tensor(f!) cnt, int M, int N, int K, "
      "int[] a_row_shift, int[] a_col_off, int[] b_row_off, int[] b_col_off, float drop_p, int seed, Tensor? seed_dev, "
      "Tensor? m_dev, int cluster, int[] gate) -> ()");
  m.def(
      "tc_gemm(Tensor A, Tensor B, Tensor(a!) out, int mode, int epi, int block_n, int M, int N, int K, "
      "int[] a_row_shift, int[] a_col_off, int[] b_row_off, int[] b_col_off, int splits, int win_w, "
      "Tensor? bias, Tensor? which, Tensor? add_src, Tensor? row_scale, Tensor? m_dev, int max_ctas, int cluster, "
      "int[] gate) -> ()");
}
void register_gemm_impls(torch::Library& m) {
  m.impl("tc_gemm", tc_gemm);
  m.impl("tc_gemm_maxout_ln", tc_gemm_maxout_ln);
}

}  // namespace srb
