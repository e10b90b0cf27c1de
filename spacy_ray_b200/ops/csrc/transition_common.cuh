// Pieces shared by the two transition kernels (ner_kernels.cu / parser_kernels.cu): vector
// loads of precomputed feature rows and the upper layer (hidden -> action scores) out of
// shared memory.
#pragma once
#include "common.cuh"

namespace srb {

// PPL consecutive bf16 (this lane's pre-activations of one feature row) -> fp32, with the
// widest loads the alignment of `lane * PPL` elements allows.
template <int PPL>
__device__ __forceinline__ void load_bf16_vec(const __nv_bfloat16* __restrict__ p, float out[PPL]) {
  if constexpr (PPL % 8 == 0) {
#pragma unroll
    for (int v = 0; v < PPL / 8; ++v) {
      const uint4 raw = *(const uint4*)(p + v * 8);
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        out[v * 8 + 2 * i] = __uint_as_float(w[i] << 16);
        out[v * 8 + 2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      }
    }
  } else if constexpr (PPL % 4 == 0) {
#pragma unroll
    for (int v = 0; v < PPL / 4; ++v) {
      const uint2 raw = *(const uint2*)(p + v * 4);
      out[v * 4 + 0] = __uint_as_float(raw.x << 16); out[v * 4 + 1] = __uint_as_float(raw.x & 0xFFFF0000u);
      out[v * 4 + 2] = __uint_as_float(raw.y << 16); out[v * 4 + 3] = __uint_as_float(raw.y & 0xFFFF0000u);
    }
  } else if constexpr (PPL % 2 == 0) {
#pragma unroll
    for (int v = 0; v < PPL / 2; ++v) {
      const uint32_t raw = *(const uint32_t*)(p + v * 2);
      out[v * 2] = __uint_as_float(raw << 16); out[v * 2 + 1] = __uint_as_float(raw & 0xFFFF0000u);
    }
  } else {
#pragma unroll
    for (int k = 0; k < PPL; ++k) out[k] = bf2f(p[k]);
  }
}

// This lane's UPL hidden activations + winning pieces of one step record, packed stores.
template <int UPL>
__device__ __forceinline__ void store_hidden_record(uint8_t* which_dst, __nv_bfloat16* hid_dst, const float best[UPL],
                                                    const uint8_t which[UPL]) {
  if constexpr (UPL == 4) {
    *(uint32_t*)which_dst = (uint32_t)which[0] | ((uint32_t)which[1] << 8) | ((uint32_t)which[2] << 16) |
                            ((uint32_t)which[3] << 24);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(best[0], best[1]), hi = __floats2bfloat162_rn(best[2], best[3]);
    uint2 v;
    v.x = *(const uint32_t*)&lo; v.y = *(const uint32_t*)&hi;
    *(uint2*)hid_dst = v;
  } else if constexpr (UPL == 2) {
    *(uint16_t*)which_dst = (uint16_t)((uint32_t)which[0] | ((uint32_t)which[1] << 8));
    *(__nv_bfloat162*)hid_dst = __floats2bfloat162_rn(best[0], best[1]);
  } else {
#pragma unroll
    for (int u = 0; u < UPL; ++u) { which_dst[u] = which[u]; hid_dst[u] = f2bf(best[u]); }
  }
}

// Upper-layer weights in shared memory as float4 per (group of 4 hidden units, action):
//   Wu4[(o / 4) * nA_pad + a] = { Wu[a][o], Wu[a][o+1], Wu[a][o+2], Wu[a][o+3] }   (fp32)
// so one LDS.128 feeds four FMAs (the [o][a] scalar layout needed one LDS per FMA and the
// kernels were issue-bound on exactly that).
__device__ __forceinline__ void stage_upper_weights(float4* Wu4, float* bu_s, const __nv_bfloat16* __restrict__ Wu,
                                                    const __nv_bfloat16* __restrict__ bu, int nO, int nA, int nA_pad) {
  float* flat = (float*)Wu4;
  for (int i = threadIdx.x; i < nO * nA_pad; i += blockDim.x) {
    // i enumerates (og, a, q): element q of the float4 at Wu4[og * nA_pad + a]
    const int q = i & 3, a = (i >> 2) % nA_pad, og = (i >> 2) / nA_pad;
    flat[i] = a < nA ? bf2f(Wu[(size_t)a * nO + og * 4 + q]) : 0.f;
  }
  for (int i = threadIdx.x; i < nA_pad; i += blockDim.x) bu_s[i] = i < nA ? bf2f(bu[i]) : 0.f;
}

// scores[a] = bu[a] + sum_o hid[o] * Wu[a][o] for this lane's actions a = lane + 32 j.
// hid_w: this warp's hidden vector in shared memory (16-byte aligned, nO % 8 == 0).
template <int NJ>
__device__ __forceinline__ void upper_layer(const float4* __restrict__ Wu4, const float* __restrict__ bu_s,
                                            const float* __restrict__ hid_w, int nO, int nA_pad, int lane,
                                            float sc[NJ]) {
  constexpr int KA = NJ == 1 ? 4 : 2;          // independent accumulation chains per action
  float acc[NJ][KA];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    acc[j][0] = bu_s[lane + 32 * j < nA_pad ? lane + 32 * j : 0];
#pragma unroll
    for (int q = 1; q < KA; ++q) acc[j][q] = 0.f;
  }
  const int n_groups = nO >> 2;
#pragma unroll 2
  for (int og = 0; og < n_groups; og += KA) {
#pragma unroll
    for (int q = 0; q < KA; ++q) {
      const float4 h = *(const float4*)(hid_w + (og + q) * 4);
      const float4* wrow = Wu4 + (size_t)(og + q) * nA_pad + lane;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (j < NJ - 1 || lane + 32 * j < nA_pad) {
          const float4 w = wrow[32 * j];
          acc[j][q] = fmaf(h.x, w.x, fmaf(h.y, w.y, fmaf(h.z, w.z, fmaf(h.w, w.w, acc[j][q]))));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float s = acc[j][0];
#pragma unroll
    for (int q = 1; q < KA; ++q) s += acc[j][q];
    sc[j] = s;
  }
}

}  // namespace srb
