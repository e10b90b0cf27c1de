// Fused non-GEMM kernels of the tok2vec / tagger hot path.
// Each one replaces a chain of thinc/cupy launches with a single pass over HBM.
#include <cstdlib>
#include "common.cuh"
#include "launch.h"
#include "gate.cuh"
#include "kernels.h"

namespace srb {

// programmatic dependent launch switch (launch.h): off unless SRB_PDL=1 - measured 1.8 % SLOWER on the
// flagship step (profiles/r2_pdl.md): the captured step keeps the GPU busy, there is no launch gap to hide
int g_pdl = []() { const char* e = getenv("SRB_PDL"); return (e && e[0] == '1') ? 1 : 0; }();


// =====================================================================================
// K1  MultiHashEmbed forward: hash -> 4-row gather-sum -> concat, one pass.
// =====================================================================================
// One WARP per row (8 rows per block): lane a computes the four table rows of attribute a once
// (2 x fmix64 + 4 runtime modulos - round 1 had every 8-column thread redo them and the kernel was
// issue-bound at 82 % with 1.9 % DRAM), the warp shares them by shuffle, then the lanes walk the
// (table, 8-column) pairs of the output row with 16-byte gathers.
__global__ void __launch_bounds__(256) hash_embed_fwd_kernel(const int64_t* __restrict__ attrs,
                                                             const float* __restrict__ mask, HashEmbedTables t,
                                                             __nv_bfloat16* __restrict__ out, int Tp,
                                                             GateArgs gate) {
  pdl_prologue();
  if (gate.flags != nullptr) {
    // C2: the embedding tables are the first weights the forward pass reads and the last the exchange
    // publishes; wait for their owners' "published" flags here instead of at the end of the last step
    if (threadIdx.x < 32) gate_wait_warp(gate);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int w8 = t.width / 8;                         // 16-byte vectors per table row
  const int pairs = t.n_tables * w8;                  // (table, vector) pairs of one output row
  const int C = t.n_tables * t.width;
  const bool pow2 = (w8 & (w8 - 1)) == 0;
  const int sh = 31 - __clz(w8);
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = gwarp; row < Tp; row += nwarps) {
    bf16x8* orow = (bf16x8*)(out + (size_t)row * C);
    if (mask[row] == 0.0f) {                          // warp-uniform
      bf16x8 z;
#pragma unroll
      for (int i = 0; i < 8; ++i) z.v[i] = f2bf(0.0f);
      for (int p = lane; p < pairs; p += 32) orow[p] = z;
      continue;
    }
    uint32_t mine[4] = {0u, 0u, 0u, 0u};
    if (lane < t.n_tables)
      hash_rows((uint64_t)attrs[(size_t)row * t.n_attr + t.column[lane]], t.seed[lane], t.n_rows[lane], mine);
    for (int p0 = 0; p0 < pairs; p0 += 32) {          // trip count is warp-uniform: the shuffles below are safe
      const int p = p0 + lane;
      const bool ok = p < pairs;
      const int a = ok ? (pow2 ? (p >> sh) : (p / w8)) : 0;
      const int within = (p - a * w8) * 8;
      uint32_t rows[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) rows[k] = __shfl_sync(0xffffffffu, mine[k], a);
      if (!ok) continue;
      const __nv_bfloat16* E = (const __nv_bfloat16*)t.table[a];
      bf16x8 r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k] = *(const bf16x8*)(E + (size_t)rows[k] * t.width + within);
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += bf2f(r[k].v[i]);
      bf16x8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = f2bf(acc[i]);
      orow[p] = o;
    }
  }
}

void launch_hash_embed_fwd(const int64_t* attrs, const float* mask, HashEmbedTables t, void* out, int Tp,
                           const GateArgs& gate, cudaStream_t s) {
  if (Tp <= 0) return;
  int blocks = (Tp + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(hash_embed_fwd_kernel, blocks, 256, 0, s, attrs, mask, t, (__nv_bfloat16*)out, Tp, gate);
}

// K1 backward: scatter-add into the fp32 table gradients.
__global__ void __launch_bounds__(256) hash_embed_bwd_kernel(const int64_t* __restrict__ attrs,
                                                             const float* __restrict__ mask, HashEmbedTables t,
                                                             const __nv_bfloat16* __restrict__ dY, int Tp) {
  pdl_prologue();
  const int row = blockIdx.x;
  if (mask[row] == 0.0f) return;
  const int C = t.n_tables * t.width;
  for (int v = threadIdx.x; v < C / 8; v += blockDim.x) {
    const int col = v * 8;
    const int a = col / t.width;
    const int within = col - a * t.width;
    uint32_t rows[4];
    hash_rows((uint64_t)attrs[(size_t)row * t.n_attr + t.column[a]], t.seed[a], t.n_rows[a], rows);
    bf16x8 g = *(const bf16x8*)(dY + (size_t)row * C + col);
    float gf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gf[i] = bf2f(g.v[i]);
    float* dE = t.grad[a];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float* dst = dE + (size_t)rows[k] * t.width + within;
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(dst + i, gf[i]);
    }
  }
}

void launch_hash_embed_bwd(const int64_t* attrs, const float* mask, HashEmbedTables t, const void* dY, int Tp,
                           cudaStream_t s) {
  if (Tp <= 0) return;
  int C = t.n_tables * t.width;
  int threads = C / 8 < 256 ? ((C / 8 + 31) / 32) * 32 : 256;
  launch_k(hash_embed_bwd_kernel, Tp, threads, 0, s, attrs, mask, t, (const __nv_bfloat16*)dY, Tp);
}

// =====================================================================================
// K2 epilogue + K3 + K5: bias, maxout, LayerNorm, dropout, residual, mask.  Warp per row.
// =====================================================================================
constexpr int kMaxUPL = 16;   // units per lane: nO <= 512

template <int NP>
__global__ void __launch_bounds__(128) maxout_ln_fwd_kernel(
    const __nv_bfloat16* __restrict__ Z, const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* __restrict__ G,
    const __nv_bfloat16* __restrict__ beta, const __nv_bfloat16* __restrict__ Xres, const float* __restrict__ mask,
    __nv_bfloat16* __restrict__ Y, uint8_t* __restrict__ which, __nv_bfloat16* __restrict__ xhat_out,
    float* __restrict__ rstd_out, int Tp, int nO, float drop_p, uint64_t seed, const int64_t* __restrict__ seed_dev) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (seed_dev) seed += (uint64_t)*seed_dev;     // device-side stream position (advances per CUDA-graph replay)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nZ = nO * NP;
  __nv_bfloat16* zs = (__nv_bfloat16*)smem_raw + (size_t)warp * nZ;
  const int upl = nO >> 5;
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const bool has_ln = G != nullptr;
  for (int row = blockIdx.x * 4 + warp; row < Tp; row += gridDim.x * 4) {
    const float m = mask[row];
    if (m == 0.0f) {          // pad row: keep everything exactly zero
      for (int u = lane; u < nO; u += 32) {
        Y[(size_t)row * nO + u] = f2bf(0.f);
        if (which) which[(size_t)row * nO + u] = 0;
        if (xhat_out) xhat_out[(size_t)row * nO + u] = f2bf(0.f);
      }
      if (lane == 0 && rstd_out) rstd_out[row] = 0.f;
      continue;
    }
    // stage the Z row (coalesced 16B loads)
    const bf16x8* zrow = (const bf16x8*)(Z + (size_t)row * nZ);
    for (int v = lane; v < nZ / 8; v += 32) ((bf16x8*)zs)[v] = zrow[v];
    __syncwarp();
    float h[kMaxUPL];
    uint8_t wh[kMaxUPL];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxUPL; ++j) {
      if (j < upl) {
        const int u = lane + 32 * j;
        float best = bf2f(zs[u * NP]) + (bias ? bf2f(bias[u * NP]) : 0.f);
        int bi = 0;
#pragma unroll
        for (int p = 1; p < NP; ++p) {
          float v = bf2f(zs[u * NP + p]) + (bias ? bf2f(bias[u * NP + p]) : 0.f);
          if (v > best) { best = v; bi = p; }
        }
        h[j] = best; wh[j] = (uint8_t)bi; sum += best;
      }
    }
    float rstd = 1.f, mu = 0.f;
    if (has_ln) {
      mu = warp_sum(sum) / (float)nO;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxUPL; ++j) if (j < upl) { float d = h[j] - mu; sq += d * d; }
      float var = warp_sum(sq) / (float)nO + 1e-8f;
      rstd = rsqrtf(var);
    }
#pragma unroll
    for (int j = 0; j < kMaxUPL; ++j) {
      if (j < upl) {
        const int u = lane + 32 * j;
        const size_t idx = (size_t)row * nO + u;
        float xh = (h[j] - mu) * rstd;
        float n = has_ln ? xh * bf2f(G[u]) + bf2f(beta[u]) : h[j];
        if (drop_p > 0.f) n *= dropout_scale(seed, idx, drop_p, inv_keep);
        if (Xres) n += bf2f(Xres[idx]);
        Y[idx] = f2bf(n);
        if (which) which[idx] = wh[j];
        if (xhat_out) xhat_out[idx] = f2bf(xh);
      }
    }
    if (lane == 0 && rstd_out) rstd_out[row] = rstd;
    __syncwarp();
  }
}

void launch_maxout_ln_fwd(const void* Z, const void* bias, const void* G, const void* beta, const void* X_res,
                          const float* mask, void* Y, uint8_t* which, void* xhat, float* rstd, int Tp, int nO,
                          int nP, float drop_p, uint64_t seed, const int64_t* seed_dev, cudaStream_t s) {
  if (Tp <= 0) return;
  if (try_launch_maxout_ln_fwd_vec(Z, bias, G, beta, X_res, mask, Y, which, xhat, rstd, Tp, nO, nP, drop_p, seed,
                                   seed_dev, s))
    return;
  int blocks = (Tp + 3) / 4;
  if (blocks > 148 * 16) blocks = 148 * 16;
  size_t smem = (size_t)4 * nO * nP * sizeof(__nv_bfloat16);
#define SRB_LAUNCH(NP)                                                                                   \
  launch_k(maxout_ln_fwd_kernel<NP>, blocks, 128, smem, s,                                                     \
      (const __nv_bfloat16*)Z, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)G, (const __nv_bfloat16*)beta, \
      (const __nv_bfloat16*)X_res, mask, (__nv_bfloat16*)Y, which, (__nv_bfloat16*)xhat, rstd, Tp, nO, drop_p, seed, \
      seed_dev)
  if (nP == 2) SRB_LAUNCH(2);
  else if (nP == 3) SRB_LAUNCH(3);
  else SRB_LAUNCH(1);
#undef SRB_LAUNCH
}

// Backward: dY -> (dropout, LN) -> dH -> routed dZ; accumulates dG, dbeta, db.
template <int NP>
__global__ void __launch_bounds__(128) maxout_ln_bwd_kernel(
    const __nv_bfloat16* __restrict__ dY, const __nv_bfloat16* __restrict__ xhat, const float* __restrict__ rstd_in,
    const __nv_bfloat16* __restrict__ G, const uint8_t* __restrict__ which, const float* __restrict__ mask,
    __nv_bfloat16* __restrict__ dZ, float* __restrict__ db, float* __restrict__ dG, float* __restrict__ dbeta,
    int Tp, int nO, float drop_p, uint64_t seed, const int64_t* __restrict__ seed_dev, int has_ln) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (seed_dev) seed += (uint64_t)*seed_dev;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nZ = nO * NP;
  __nv_bfloat16* zs = (__nv_bfloat16*)smem_raw + (size_t)warp * nZ;
  const int upl = nO >> 5;
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  float accG[kMaxUPL], accB[kMaxUPL], accb[kMaxUPL * NP];
#pragma unroll
  for (int j = 0; j < kMaxUPL; ++j) { accG[j] = 0.f; accB[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < kMaxUPL * NP; ++j) accb[j] = 0.f;
  for (int row = blockIdx.x * 4 + warp; row < Tp; row += gridDim.x * 4) {
    bf16x8* zout = (bf16x8*)(dZ + (size_t)row * nZ);
    if (mask[row] == 0.0f) {
      bf16x8 zero;
#pragma unroll
      for (int i = 0; i < 8; ++i) zero.v[i] = f2bf(0.f);
      for (int v = lane; v < nZ / 8; v += 32) zout[v] = zero;
      continue;
    }
    const float rstd = has_ln ? rstd_in[row] : 1.f;
    float dn[kMaxUPL], xh[kMaxUPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxUPL; ++j) {
      if (j < upl) {
        const int u = lane + 32 * j;
        const size_t idx = (size_t)row * nO + u;
        float d = bf2f(dY[idx]);
        if (drop_p > 0.f) d *= dropout_scale(seed, idx, drop_p, inv_keep);
        dn[j] = d;
        if (has_ln) {
          xh[j] = bf2f(xhat[idx]);
          accG[j] += d * xh[j];
          accB[j] += d;
          float dx = d * bf2f(G[u]);
          dn[j] = dx;
          s1 += dx; s2 += dx * xh[j];
        }
      }
    }
    if (has_ln) {
      s1 = warp_sum(s1) / (float)nO;
      s2 = warp_sum(s2) / (float)nO;
    }
    for (int v = lane; v < nZ / 8; v += 32) {
      bf16x8 zero;
#pragma unroll
      for (int i = 0; i < 8; ++i) zero.v[i] = f2bf(0.f);
      ((bf16x8*)zs)[v] = zero;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kMaxUPL; ++j) {
      if (j < upl) {
        const int u = lane + 32 * j;
        float dH = has_ln ? rstd * (dn[j] - s1 - xh[j] * s2) : dn[j];
        const int p = which[(size_t)row * nO + u];
        zs[u * NP + p] = f2bf(dH);
#pragma unroll
        for (int q = 0; q < NP; ++q) accb[j * NP + q] += (q == p) ? dH : 0.f;
      }
    }
    __syncwarp();
    for (int v = lane; v < nZ / 8; v += 32) zout[v] = ((bf16x8*)zs)[v];
    __syncwarp();
  }
  // block-level combine of the per-warp register accumulators, then one atomic per element
  __syncthreads();
  float* red = (float*)smem_raw;       // reuse: needs 4 * nO * (NP + 2) floats
  float* rG = red, *rB = red + 4 * nO, *rb = red + 8 * nO;
#pragma unroll
  for (int j = 0; j < kMaxUPL; ++j) {
    if (j < upl) {
      const int u = lane + 32 * j;
      rG[warp * nO + u] = accG[j];
      rB[warp * nO + u] = accB[j];
#pragma unroll
      for (int q = 0; q < NP; ++q) rb[warp * nZ + u * NP + q] = accb[j * NP + q];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nO; i += blockDim.x) {
    if (has_ln) {
      atomicAdd(dG + i, rG[i] + rG[nO + i] + rG[2 * nO + i] + rG[3 * nO + i]);
      atomicAdd(dbeta + i, rB[i] + rB[nO + i] + rB[2 * nO + i] + rB[3 * nO + i]);
    }
  }
  for (int i = threadIdx.x; i < nZ; i += blockDim.x)
    atomicAdd(db + i, rb[i] + rb[nZ + i] + rb[2 * nZ + i] + rb[3 * nZ + i]);
}

void launch_maxout_ln_bwd(const void* dY, const void* xhat, const float* rstd, const void* G, const uint8_t* which,
                          const float* mask, void* dZ, float* db, float* dG, float* dbeta, int Tp, int nO, int nP,
                          float drop_p, uint64_t seed, const int64_t* seed_dev, int has_ln, cudaStream_t s) {
  if (Tp <= 0) return;
  if (try_launch_maxout_ln_bwd_vec(dY, xhat, rstd, G, which, mask, dZ, db, dG, dbeta, Tp, nO, nP, drop_p, seed,
                                   seed_dev, has_ln, s))
    return;
  int blocks = (Tp + 3) / 4;
  if (blocks > 148 * 4) blocks = 148 * 4;
  size_t smem_z = (size_t)4 * nO * nP * sizeof(__nv_bfloat16);
  size_t smem_r = (size_t)4 * nO * (nP + 2) * sizeof(float);
  size_t smem = smem_z > smem_r ? smem_z : smem_r;
#define SRB_LAUNCH(NP)                                                                                          \
  do {                                                                                                          \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(maxout_ln_bwd_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    launch_k(maxout_ln_bwd_kernel<NP>, blocks, 128, smem, s,                                                          \
        (const __nv_bfloat16*)dY, (const __nv_bfloat16*)xhat, rstd, (const __nv_bfloat16*)G, which, mask,       \
        (__nv_bfloat16*)dZ, db, dG, dbeta, Tp, nO, drop_p, seed, seed_dev, has_ln);                                       \
  } while (0)
  if (nP == 2) SRB_LAUNCH(2);
  else if (nP == 3) SRB_LAUNCH(3);
  else SRB_LAUNCH(1);
#undef SRB_LAUNCH
}

// =====================================================================================
// K4 (library-GEMM path only): materialised window and its adjoint.
// =====================================================================================
__global__ void seq2col_kernel(const bf16x8* __restrict__ X, bf16x8* __restrict__ Xw, int Tp, int nI8) {
  pdl_prologue();
  const size_t total = (size_t)Tp * 3 * nI8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (3 * nI8));
    const int t = (int)(i / (3 * nI8));
    const int blk = c / nI8, within = c - blk * nI8;
    const int src = t + blk - 1;
    bf16x8 v;
    if (src >= 0 && src < Tp) v = X[(size_t)src * nI8 + within];
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] = f2bf(0.f);
    }
    Xw[i] = v;
  }
}
void launch_seq2col(const void* X, void* Xw, int Tp, int nI, cudaStream_t s) {
  if (Tp <= 0) return;
  size_t total = (size_t)Tp * 3 * (nI / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  launch_k(seq2col_kernel, blocks, 256, 0, s, (const bf16x8*)X, (bf16x8*)Xw, Tp, nI / 8);
}

__global__ void col2seq_residual_kernel(const bf16x8* __restrict__ dXw, const bf16x8* __restrict__ dY,
                                        const float* __restrict__ mask, bf16x8* __restrict__ dX, int Tp, int nI8,
                                        int add_res) {
  pdl_prologue();
  const size_t total = (size_t)Tp * nI8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % nI8);
    const int t = (int)(i / nI8);
    float acc[8];
    bf16x8 v = dXw[(size_t)t * 3 * nI8 + nI8 + c];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = bf2f(v.v[k]);
    if (t + 1 < Tp) {
      v = dXw[(size_t)(t + 1) * 3 * nI8 + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += bf2f(v.v[k]);
    }
    if (t >= 1) {
      v = dXw[(size_t)(t - 1) * 3 * nI8 + 2 * nI8 + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += bf2f(v.v[k]);
    }
    if (add_res) {
      const float m = mask[t];
      v = dY[i];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += m * bf2f(v.v[k]);
    }
    bf16x8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = f2bf(acc[k]);
    dX[i] = o;
  }
}
void launch_col2seq_residual(const void* dXw, const void* dY, const float* mask, void* dX, int Tp, int nI,
                             int add_residual, cudaStream_t s) {
  if (Tp <= 0) return;
  size_t total = (size_t)Tp * (nI / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  launch_k(col2seq_residual_kernel, blocks, 256, 0, s, (const bf16x8*)dXw, (const bf16x8*)dY, mask, (bf16x8*)dX, Tp,
                                                 nI / 8, add_residual);
}

// =====================================================================================
// K6  softmax + cross-entropy gradient + loss + argmax.  Warp per row, nC <= 256.
// =====================================================================================
__global__ void __launch_bounds__(128) softmax_xent_kernel(const float* __restrict__ logits,
                                                           const int64_t* __restrict__ labels,
                                                           __nv_bfloat16* __restrict__ d_out,
                                                           int64_t* __restrict__ guesses, float* __restrict__ loss,
                                                           int Tp, int nC) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float local_loss = 0.f;
  for (int row = blockIdx.x * 4 + warp; row < Tp; row += gridDim.x * 4) {
    const float* x = logits + (size_t)row * nC;
    float v[8];
    float mx = -3.0e38f;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      v[j] = c < nC ? x[c] : -3.0e38f;
      if (v[j] > mx) { mx = v[j]; arg = c; }
    }
    // warp argmax (lowest index wins ties, like torch.argmax on first max)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, mx, o);
      int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      v[j] = c < nC ? __expf(v[j] - mx) : 0.f;
      sum += v[j];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const int64_t lab = labels[row];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      if (c < nC) {
        float d = 0.f;
        if (lab >= 0) { d = v[j] * inv - (c == (int)lab ? 1.f : 0.f); local_loss += d * d; }
        d_out[(size_t)row * nC + c] = f2bf(d);
      }
    }
    if (lane == 0) guesses[row] = arg;
  }
  local_loss = warp_sum(local_loss);
  if (lane == 0 && local_loss != 0.f) atomicAdd(loss, local_loss);
}
void launch_softmax_xent(const float* logits, const int64_t* labels, void* d_out, int64_t* guesses, float* loss,
                         int Tp, int nC, cudaStream_t s) {
  if (Tp <= 0) return;
  int blocks = (Tp + 3) / 4;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(softmax_xent_kernel, blocks, 128, 0, s, logits, labels, (__nv_bfloat16*)d_out, guesses, loss, Tp, nC);
}

// =====================================================================================
// K8  multi-tensor Adam with per-tensor clipping (thinc semantics), two passes.
// =====================================================================================
// hyper = {lr, beta1, beta2, eps, grad_clip, l2, l2_is_wd, grad_scale}
__global__ void __launch_bounds__(256) adam_sumsq_kernel(const float* __restrict__ g, const int32_t* __restrict__ blk_key,
                                                         const int32_t* __restrict__ blk_off,
                                                         const int64_t* __restrict__ key_off,
                                                         const int64_t* __restrict__ key_len,
                                                         float* __restrict__ norms_sq, const float* __restrict__ hyper,
                                                         const float* __restrict__ w) {
  const int k = blk_key[blockIdx.x];
  const int64_t base = key_off[k] + (int64_t)blk_off[blockIdx.x] * kAdamChunk;
  const int64_t end = key_off[k] + key_len[k];
  const float gs = hyper[7], l2 = hyper[5];
  const bool l2_in_grad = l2 != 0.f && hyper[6] == 0.f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < kAdamChunk; i += 256) {
    const int64_t idx = base + i;
    if (idx < end) {
      float v = g[idx] * gs;
      if (l2_in_grad) v += l2 * w[idx];
      acc += v * v;
    }
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(norms_sq + k, t);
  }
}
void launch_adam_sumsq(const float* g, const int32_t* blk_key, const int32_t* blk_off, const int64_t* key_off,
                       const int64_t* key_len, float* norms_sq, int n_blocks, const float* hyper, const float* w,
                       cudaStream_t s) {
  if (n_blocks <= 0) return;
  adam_sumsq_kernel<<<n_blocks, 256, 0, s>>>(g, blk_key, blk_off, key_off, key_len, norms_sq, hyper, w);
}

__global__ void __launch_bounds__(256) adam_update_kernel(float* __restrict__ g, float* __restrict__ w,
                                                          float* __restrict__ m1, float* __restrict__ m2,
                                                          __nv_bfloat16* __restrict__ w_out,
                                                          const int32_t* __restrict__ blk_key,
                                                          const int32_t* __restrict__ blk_off,
                                                          const int64_t* __restrict__ key_off,
                                                          const int64_t* __restrict__ key_len,
                                                          const float* __restrict__ norms_sq,
                                                          const float* __restrict__ hyper,
                                                          const int32_t* __restrict__ step) {
  const int k = blk_key[blockIdx.x];
  const int64_t base = key_off[k] + (int64_t)blk_off[blockIdx.x] * kAdamChunk;
  const int64_t end = key_off[k] + key_len[k];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], clip = hyper[4], l2 = hyper[5];
  const bool wd = hyper[6] != 0.f;
  const float gs = hyper[7];
  const float t = (float)(*step + 1);
  const float fix1 = 1.f - powf(b1, t), fix2 = 1.f - powf(b2, t);
  const float lr_t = lr * sqrtf(fix2) / fix1;
  float scale = gs;
  if (clip > 0.f) {
    const float norm = sqrtf(norms_sq[k]);
    if (norm >= clip) scale *= clip / fmaxf(norm, 1e-30f);
  }
  const bool l2_in_grad = l2 != 0.f && !wd;
  for (int i = threadIdx.x; i < kAdamChunk; i += 256) {
    const int64_t idx = base + i;
    if (idx < end) {
      float wv = w[idx];
      float gv = g[idx];
      if (l2_in_grad) gv = (gv * gs + l2 * wv) * (scale / gs); else gv *= scale;
      float a = b1 * m1[idx] + (1.f - b1) * gv;
      float b = b2 * m2[idx] + (1.f - b2) * gv * gv;
      m1[idx] = a; m2[idx] = b;
      wv -= lr_t * a / (sqrtf(b) + eps);
      if (wd && l2 != 0.f) wv *= (1.f - lr * l2);
      w[idx] = wv;
      if (w_out) w_out[idx] = f2bf(wv);
      g[idx] = 0.f;
    }
  }
}
void launch_adam_update(float* g, float* w, float* m1, float* m2, void* w_out_bf16, const int32_t* blk_key,
                        const int32_t* blk_off, const int64_t* key_off, const int64_t* key_len,
                        const float* norms_sq, int n_blocks, const float* hyper, const int32_t* step,
                        cudaStream_t s) {
  if (n_blocks <= 0) return;
  adam_update_kernel<<<n_blocks, 256, 0, s>>>(g, w, m1, m2, (__nv_bfloat16*)w_out_bf16, blk_key, blk_off, key_off,
                                             key_len, norms_sq, hyper, step);
}

}  // namespace srb
