// LayerNorm -> dropout -> (+residual) -> mask for rows of already max-ed activations, one warp per
// row, a lane owning UPL contiguous units: the row math of maxout_ln_fwd_vec_kernel<1, UPL, R>
// (elementwise_fast.cu) as a device function, used by the LN warps of the fused GEMM
// (gemm_tcgen05.cu, EPI_MAXOUT3_LN).  H is read with ld.global.cg: it was written by other SMs
// during the same kernel, so the (non-coherent) L1 must be bypassed.
#pragma once
#include "common.cuh"

namespace srb {

template <int UPL, int R>
__device__ __forceinline__ void ln_fwd_rows(const __nv_bfloat16* __restrict__ H, const __nv_bfloat16* __restrict__ Xres,
                                            const float* __restrict__ mask, const float (&gk)[UPL],
                                            const float (&bk)[UPL], bool has_ln, __nv_bfloat16* __restrict__ Y,
                                            __nv_bfloat16* __restrict__ xhat_out, float* __restrict__ rstd_out,
                                            int row0, int n_rows, int Tp, int lane, float drop_p, uint32_t thr,
                                            float inv_keep, uint64_t seed) {
  constexpr int nO = 32 * UPL;
  const int u0 = lane * UPL;
  for (int base = 0; base < n_rows; base += R) {
    uint4 hraw[R][UPL / 8];
    bf16x8 xraw[R][UPL / 8];
    float mk[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = min(row0 + base + r, Tp - 1);
      mk[r] = mask[row];
#pragma unroll
      for (int v = 0; v < UPL / 8; ++v) {
        hraw[r][v] = __ldcg((const uint4*)(H + (size_t)row * nO + u0) + v);
        if (Xres) xraw[r][v] = *((const bf16x8*)(Xres + (size_t)row * nO + u0) + v);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + base + r;
      if (base + r >= n_rows || row >= Tp) break;
      const size_t ro = (size_t)row * nO + u0;
      if (mk[r] == 0.0f) {
        bf16x8 zero;
#pragma unroll
        for (int i = 0; i < 8; ++i) zero.v[i] = f2bf(0.f);
#pragma unroll
        for (int v = 0; v < UPL / 8; ++v) {
          *(bf16x8*)(Y + ro + v * 8) = zero;
          if (xhat_out) *(bf16x8*)(xhat_out + ro + v * 8) = zero;
        }
        if (lane == 0 && rstd_out) rstd_out[row] = 0.f;
        continue;
      }
      float h[UPL];
      float sum = 0.f;
#pragma unroll
      for (int v = 0; v < UPL / 8; ++v) {
        const uint32_t w[4] = {hraw[r][v].x, hraw[r][v].y, hraw[r][v].z, hraw[r][v].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h[v * 8 + 2 * i] = __uint_as_float(w[i] << 16);
          h[v * 8 + 2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
      }
#pragma unroll
      for (int j = 0; j < UPL; ++j) sum += h[j];
      float mu = 0.f, rstd = 1.f;
      if (has_ln) {
        mu = warp_sum(sum) * (1.f / nO);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < UPL; ++j) { const float d = h[j] - mu; sq += d * d; }
        rstd = rsqrtf(warp_sum(sq) * (1.f / nO) + 1e-8f);
      }
#pragma unroll
      for (int v = 0; v < UPL / 8; ++v) {
        bf16x8 yo, xo;
        float keep[8];
        if (drop_p > 0.f) dropout_scale8(seed, ro + v * 8, thr, inv_keep, keep);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = v * 8 + i;
          const float xh = (h[j] - mu) * rstd;
          float n = has_ln ? xh * gk[j] + bk[j] : h[j];
          if (drop_p > 0.f) n *= keep[i];
          if (Xres) n += bf2f(xraw[r][v].v[i]);
          yo.v[i] = f2bf(n);
          xo.v[i] = f2bf(xh);
        }
        *(bf16x8*)(Y + ro + v * 8) = yo;
        if (xhat_out) *(bf16x8*)(xhat_out + ro + v * 8) = xo;
      }
      if (lane == 0 && rstd_out) rstd_out[row] = rstd;
    }
  }
}

}  // namespace srb
