// K6: the tagger head in ONE kernel - small-N GEMM (logits = X W^T + b) + row softmax + cross-
// entropy gradient + loss + arg-max.  n_classes is tens, so this is a mat-vec per token against
// a weight matrix that lives in shared memory (same float4 layout / LDS.128-per-4-FMA scheme as
// the transition kernels' upper layer): one warp per row, a lane owns classes lane, lane+32, ...
// Upstream: cuBLAS GEMM + bias add + softmax + (p - onehot) as separate launches.
// d is written with a 128-multiple pitch (zero past n_classes) so dW = d^T X and dX = d W can go
// straight to the tcgen05 GEMMs.
#include "common.cuh"
#include "launch.h"
#include "kernels.h"
#include "transition_common.cuh"

namespace srb {

constexpr int kTagWarps = 4;

template <int NJ>
__global__ void __launch_bounds__(kTagWarps * 32) linear_softmax_xent_kernel(
    const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ b,
    const int64_t* __restrict__ labels, __nv_bfloat16* __restrict__ d_out, int64_t* __restrict__ guesses,
    float* __restrict__ loss, int Tp, int w, int nC, int nC_pad, int ldd) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* W4 = (float4*)smem_raw;                                   // [w/4][nC_pad] x float4
  float* b_s = (float*)smem_raw + (size_t)w * nC_pad;               // [nC_pad]
  float* x_s = b_s + nC_pad;                                        // [warps][w]
  stage_upper_weights(W4, b_s, W, b, w, nC, nC_pad);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xw = x_s + warp * w;
  float local_loss = 0.f;
  for (int row = blockIdx.x * kTagWarps + warp; row < Tp; row += gridDim.x * kTagWarps) {
    for (int k = lane; k < w; k += 32) xw[k] = bf2f(X[(size_t)row * w + k]);
    __syncwarp();
    float sc[NJ];
    upper_layer<NJ>(W4, b_s, xw, w, nC_pad, lane, sc);
    float mx = -3.0e38f;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      if (c >= nC) sc[j] = -3.0e38f;
      if (sc[j] > mx) { mx = sc[j]; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                               // lowest index wins ties
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float e[NJ], sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { e[j] = (lane + 32 * j) < nC ? __expf(sc[j] - mx) : 0.f; sum += e[j]; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const int64_t lab = labels[row];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      if (c < nC) {
        float d = 0.f;
        if (lab >= 0) { d = e[j] * inv - (c == (int)lab ? 1.f : 0.f); local_loss += d * d; }
        d_out[(size_t)row * ldd + c] = f2bf(d);
      }
    }
    if (lane == 0) guesses[row] = arg;
    __syncwarp();                                                    // xw is rewritten next iteration
  }
  local_loss = warp_sum(local_loss);
  if (lane == 0 && local_loss != 0.f) atomicAdd(loss, local_loss);
}

bool try_launch_linear_softmax_xent(const void* X, const void* W, const void* b, const int64_t* labels, void* d_out,
                                    int64_t* guesses, float* loss, int Tp, int w, int nC, int ldd, cudaStream_t s) {
  if (Tp <= 0) return true;
  const int nC_pad = (nC + 7) / 8 * 8;
  const size_t smem = sizeof(float) * ((size_t)w * nC_pad + nC_pad + (size_t)kTagWarps * w);
  if (w % 16 != 0 || nC > 128 || smem > 200 * 1024) return false;
  int blocks = (Tp + kTagWarps - 1) / kTagWarps;
  if (blocks > 148 * 4) blocks = 148 * 4;
  const int nj = (nC_pad + 31) / 32;
#define SRB_TAG(NJ_)                                                                                              \
  if (nj == NJ_) {                                                                                                \
    if (smem > 48 * 1024)                                                                                         \
      cudaFuncSetAttribute(linear_softmax_xent_kernel<NJ_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    launch_k(linear_softmax_xent_kernel<NJ_>, blocks, kTagWarps * 32, smem, s,                                           \
        (const __nv_bfloat16*)X, (const __nv_bfloat16*)W, (const __nv_bfloat16*)b, labels, (__nv_bfloat16*)d_out,   \
        guesses, loss, Tp, w, nC, nC_pad, ldd);                                                                   \
    return true;                                                                                                  \
  }
  SRB_TAG(1) SRB_TAG(2) SRB_TAG(3) SRB_TAG(4)
#undef SRB_TAG
  return false;
}

// ------------------------------------------------------------------------------------------
// The same head with the logits on the tensor cores: the tcgen05 GEMM accumulates X W^T into a
// zeroed fp32 scratch (ldl-pitch, EPI_ATOMIC_F32), this kernel adds the bias and does softmax +
// cross-entropy gradient + loss + arg-max, one warp per row, and leaves the scratch ZEROED for the
// next step.  The one-kernel version above keeps W in shared memory as fp32 (w x nC x 4 B): at
// width 512 that is one 4-warp block per SM and 277 us per call; this pair is ~15 us.
// ------------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(256) softmax_xent_bias_kernel(float* __restrict__ logits, const __nv_bfloat16* __restrict__ b,
                                                                const int64_t* __restrict__ labels,
                                                                __nv_bfloat16* __restrict__ d_out, int64_t* __restrict__ guesses,
                                                                float* __restrict__ loss, int Tp, int nC, int ldl, int ldd) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  float bj[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) bj[j] = (lane + 32 * j) < nC ? bf2f(b[lane + 32 * j]) : 0.f;
  float local_loss = 0.f;
  for (int row = gwarp; row < Tp; row += nwarps) {
    float* lr = logits + (size_t)row * ldl;
    float sc[NJ];
    float mx = -3.0e38f;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      sc[j] = -3.0e38f;
      if (c < nC) { sc[j] = lr[c] + bj[j]; lr[c] = 0.f; }
      if (sc[j] > mx) { mx = sc[j]; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                               // lowest index wins ties
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float e[NJ], sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { e[j] = (lane + 32 * j) < nC ? __expf(sc[j] - mx) : 0.f; sum += e[j]; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const int64_t lab = labels[row];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      if (c < nC) {
        float d = 0.f;
        if (lab >= 0) { d = e[j] * inv - (c == (int)lab ? 1.f : 0.f); local_loss += d * d; }
        d_out[(size_t)row * ldd + c] = f2bf(d);
      }
    }
    if (lane == 0) guesses[row] = arg;
  }
  // one atomic per block (thousands of warps on one address serialise: 28 us of a 36 us kernel)
  __shared__ float wsum[8];
  local_loss = warp_sum(local_loss);
  if (lane == 0) wsum[threadIdx.x >> 5] = local_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += wsum[i];
    if (t != 0.f) atomicAdd(loss, t);
  }
}

bool launch_softmax_xent_bias(float* logits, const void* b, const int64_t* labels, void* d_out, int64_t* guesses,
                              float* loss, int Tp, int nC, int ldl, int ldd, cudaStream_t s) {
  if (Tp <= 0) return true;
  if (nC > 128 || nC > ldl) return false;
  int blocks = (Tp + 31) / 32;                       // >= 4 rows per warp
  if (blocks > 148 * 4) blocks = 148 * 4;
  const int nj = (nC + 31) / 32;
#define SRB_SM(NJ_)                                                                                              \
  if (nj == NJ_) {                                                                                               \
    launch_k(softmax_xent_bias_kernel<NJ_>, blocks, 256, 0, s, logits, (const __nv_bfloat16*)b, labels,           \
             (__nv_bfloat16*)d_out, guesses, loss, Tp, nC, ldl, ldd);                                            \
    return true;                                                                                                 \
  }
  SRB_SM(1) SRB_SM(2) SRB_SM(3) SRB_SM(4)
#undef SRB_SM
  return false;
}

}  // namespace srb
