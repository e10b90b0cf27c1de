// Vectorised fast paths for the memory-bound kernels (profiles/r1_run2_launch_list.md showed
// the scalar versions at 16x their HBM roofline):
//   * maxout_ln fwd / bwd: each lane owns UPL *contiguous* units, so every global access is
//     a 16-byte vector (no shared-memory staging), one warp per row, many warps per SM.
//   * hash_embed backward over *sorted* attribute ids: a warp walks a chunk of the sorted
//     order, sums the dY rows of a run of identical ids in registers and issues the four
//     table-row updates once per run (vector RED).  The unsorted kernel hammered ~50 hot
//     PREFIX/SHAPE rows with one atomic per token per element.
#include "common.cuh"
#include "launch.h"
#include "kernels.h"

namespace srb {

// ------------------------------------------------------------------------------------------
// forward: Z (Tp, nO*NP) [+bias] -> maxout -> LN -> dropout -> (+X) -> mask
//
// A lane owns UPL contiguous units of a row (16-byte vectors everywhere); a row takes nO / UPL
// lanes of a SEG-lane segment (SEG = 32: one row per warp; SEG = 16: two rows per warp, for widths
// up to 128 - round 1 had fast kernels for nO = 32 * UPL only, so width 96 - BASELINE config 1's
// model - ran the scalar kernels: 62 us per backward call, half of that step).  Lanes past the row
// idle and contribute zeros to the segment reductions.  A row is only 16-48 B per lane, so a warp
// walks R row groups per iteration and issues every load of all of them before the first use:
// that is what keeps enough bytes in flight per SM to cover HBM latency (the one-row version sat
// at ~35% of copy bandwidth with 24 resident warps x 1 KB).
// ------------------------------------------------------------------------------------------
template <int SEG>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
  for (int o = SEG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// FULL: the row fills its segment exactly (nO == SEG * UPL: 256, 512, 128) - nO is a compile-time
// constant and no lane idles.  With one row per warp (SEG == 32) pad rows and rows past the end are
// warp-uniform and leave early; with two rows per warp they are handled after the reductions.
template <int NP, int UPL, int R, int SEG, bool FULL>
__global__ void __launch_bounds__(256, 2) maxout_ln_fwd_vec_kernel(
    const __nv_bfloat16* __restrict__ Z, const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* __restrict__ G,
    const __nv_bfloat16* __restrict__ beta, const __nv_bfloat16* __restrict__ Xres, const float* __restrict__ mask,
    __nv_bfloat16* __restrict__ Y, uint8_t* __restrict__ which, __nv_bfloat16* __restrict__ xhat_out,
    float* __restrict__ rstd_out, int Tp, int nO_rt, float drop_p, uint64_t seed, const int64_t* __restrict__ seed_dev) {
  pdl_prologue();
  constexpr int ZPL = UPL * NP;                 // Z elements per lane (multiple of 8)
  constexpr int RPW = 32 / SEG;                 // rows per warp and r
  constexpr bool UNIFORM = SEG == 32;
  const int nO = FULL ? SEG * UPL : nO_rt;
  if (seed_dev) seed += (uint64_t)*seed_dev;
  const int lane = threadIdx.x & 31;
  const int seg = lane / SEG, sl = lane % SEG;
  const bool active = FULL ? true : (sl * UPL < nO);
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const float inv_n = 1.f / (float)nO;
  const uint32_t thr = dropout_thr(drop_p);
  const bool has_ln = G != nullptr;
  const int u0 = active ? sl * UPL : 0;
  float gk[UPL], bk[UPL];
#pragma unroll
  for (int j = 0; j < UPL; ++j) { gk[j] = has_ln ? bf2f(G[u0 + j]) : 1.f; bk[j] = has_ln ? bf2f(beta[u0 + j]) : 0.f; }
  const int ngroups = (Tp + R * RPW - 1) / (R * RPW);
  for (int grp = gwarp; grp < ngroups; grp += nwarps) {
    const int row0 = grp * R * RPW + seg;
    // ---- phase 1: every load of the R rows (pad rows are valid memory, loaded unconditionally)
    bf16x8 zraw[R][ZPL / 8];
    bf16x8 xraw[R][UPL / 8];
    float mk[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = min(row0 + r * RPW, Tp - 1);
      mk[r] = mask[row];
      const bf16x8* zp = (const bf16x8*)(Z + (size_t)row * nO * NP + u0 * NP);
#pragma unroll
      for (int v = 0; v < ZPL / 8; ++v) zraw[r][v] = zp[v];
      if (Xres) {
        const bf16x8* xp = (const bf16x8*)(Xres + (size_t)row * nO + u0);
#pragma unroll
        for (int v = 0; v < UPL / 8; ++v) xraw[r][v] = xp[v];
      }
    }
    // ---- phase 2: one row at a time out of registers
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * RPW;
      const bool row_ok = row < Tp;
      if (UNIFORM && !row_ok) break;
      const size_t ro = (size_t)(UNIFORM ? row : min(row, Tp - 1)) * nO + u0;
      const bool pad = mk[r] == 0.0f;
      auto store_zeros = [&]() {
        bf16x8 zero;
#pragma unroll
        for (int i = 0; i < 8; ++i) zero.v[i] = f2bf(0.f);
#pragma unroll
        for (int v = 0; v < UPL / 8; ++v) {
          *(bf16x8*)(Y + ro + v * 8) = zero;
          if (xhat_out) *(bf16x8*)(xhat_out + ro + v * 8) = zero;
          if (which) *(uint2*)(which + ro + v * 8) = make_uint2(0u, 0u);
        }
        if (sl == 0 && rstd_out) rstd_out[row] = 0.f;
      };
      if (UNIFORM && pad) {
        if (active) store_zeros();
        continue;
      }
      float h[UPL];
      uint8_t wh[UPL];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < UPL; ++j) {
        float best = 0.f;
        int bi = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int e = j * NP + p;
          float zv = bf2f(zraw[r][e / 8].v[e % 8]);
          if (bias) zv += bf2f(bias[u0 * NP + e]);
          if (p == 0 || zv > best) { best = zv; bi = p; }
        }
        if (!FULL && !active) best = 0.f;
        h[j] = best; wh[j] = (uint8_t)bi; sum += best;
      }
      float mu = 0.f, rstd = 1.f;
      if (has_ln) {
        mu = seg_sum<SEG>(sum) * inv_n;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < UPL; ++j) { const float d = h[j] - mu; sq += d * d; }
        if (!FULL && !active) sq = 0.f;
        rstd = rsqrtf(seg_sum<SEG>(sq) * inv_n + 1e-8f);
      }
      if (!active) continue;
      if (!UNIFORM) {                            // two rows per warp: only now may the halves diverge
        if (!row_ok) continue;
        if (pad) { store_zeros(); continue; }
      }
#pragma unroll
      for (int v = 0; v < UPL / 8; ++v) {
        bf16x8 yo, xo;
        __align__(8) uint8_t w8[8];
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
          w8[i] = wh[j];
        }
        *(bf16x8*)(Y + ro + v * 8) = yo;
        if (xhat_out) *(bf16x8*)(xhat_out + ro + v * 8) = xo;
        if (which) *(uint2*)(which + ro + v * 8) = *(const uint2*)w8;
      }
      if (sl == 0 && rstd_out) rstd_out[row] = rstd;
    }
  }
}

// rows per warp iteration: aim at ~128 B of loads in flight per lane
constexpr int rows_per_iter(int bytes_per_lane_row) {
  return bytes_per_lane_row <= 32 ? 4 : (bytes_per_lane_row <= 64 ? 2 : 1);
}

template <int NP, int UPL, int SEG, bool FULL>
static void launch_fwd_vec(const void* Z, const void* bias, const void* G, const void* beta, const void* X_res,
                           const float* mask, void* Y, uint8_t* which, void* xhat, float* rstd, int Tp, int nO,
                           float drop_p, uint64_t seed, const int64_t* seed_dev, cudaStream_t s) {
  constexpr int R = rows_per_iter((UPL * NP + UPL) * 2);
  constexpr int RPB = 8 * R * (32 / SEG);             // rows per block and iteration
  int blocks = (Tp + RPB - 1) / RPB;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_k(maxout_ln_fwd_vec_kernel<NP, UPL, R, SEG, FULL>, blocks, 256, 0, s,
      (const __nv_bfloat16*)Z, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)G, (const __nv_bfloat16*)beta,
      (const __nv_bfloat16*)X_res, mask, (__nv_bfloat16*)Y, which, (__nv_bfloat16*)xhat, rstd, Tp, nO, drop_p, seed,
      seed_dev);
}

// lane layout for a width: 8 units per lane up to 256 units (16-lane segments up to 128), 16 per lane up to 512
static bool vec_layout(int nO, int* upl, int* seg) {
  if (nO <= 0 || nO % 8) return false;
  if (nO <= 256) { *upl = 8; *seg = nO <= 128 ? 16 : 32; return true; }
  if (nO <= 512 && nO % 16 == 0) { *upl = 16; *seg = 32; return true; }
  return false;
}

bool try_launch_maxout_ln_fwd_vec(const void* Z, const void* bias, const void* G, const void* beta, const void* X_res,
                                  const float* mask, void* Y, uint8_t* which, void* xhat, float* rstd, int Tp, int nO,
                                  int nP, float drop_p, uint64_t seed, const int64_t* seed_dev, cudaStream_t s) {
  int upl, seg;
  if (!vec_layout(nO, &upl, &seg)) return false;
#define SRB_TRY(NP_, UPL_, SEG_)                                                                                   \
  if (nP == NP_ && upl == UPL_ && seg == SEG_) {                                                                   \
    if (nO == UPL_ * SEG_)                                                                                         \
      launch_fwd_vec<NP_, UPL_, SEG_, true>(Z, bias, G, beta, X_res, mask, Y, which, xhat, rstd, Tp, nO, drop_p, seed, seed_dev, s); \
    else                                                                                                           \
      launch_fwd_vec<NP_, UPL_, SEG_, false>(Z, bias, G, beta, X_res, mask, Y, which, xhat, rstd, Tp, nO, drop_p, seed, seed_dev, s); \
    return true;                                                                                                   \
  }
  SRB_TRY(1, 8, 32) SRB_TRY(1, 8, 16) SRB_TRY(1, 16, 32)
  SRB_TRY(3, 8, 32) SRB_TRY(3, 8, 16) SRB_TRY(3, 16, 32)
  SRB_TRY(2, 8, 32) SRB_TRY(2, 8, 16) SRB_TRY(2, 16, 32)
#undef SRB_TRY
  return false;
}

// ------------------------------------------------------------------------------------------
// backward: dY -> dropout -> LN backward -> routed dZ; accumulates dG, dbeta, db
// (same lane layout and R-rows-per-iteration load batching as the forward kernel)
// ------------------------------------------------------------------------------------------
template <int NP, int UPL, int R, int SEG, bool FULL>
__global__ void __launch_bounds__(256, (UPL > 8 ? 1 : 2)) maxout_ln_bwd_vec_kernel(
    const __nv_bfloat16* __restrict__ dY, const __nv_bfloat16* __restrict__ xhat, const float* __restrict__ rstd_in,
    const __nv_bfloat16* __restrict__ G, const uint8_t* __restrict__ which, const float* __restrict__ mask,
    __nv_bfloat16* __restrict__ dZ, float* __restrict__ db, float* __restrict__ dG, float* __restrict__ dbeta, int Tp,
    int nO_rt, float drop_p, uint64_t seed, const int64_t* __restrict__ seed_dev, int has_ln) {
  pdl_prologue();
  constexpr int ZPL = UPL * NP;
  constexpr int RPW = 32 / SEG;
  constexpr bool UNIFORM = SEG == 32;
  const int nO = FULL ? SEG * UPL : nO_rt;
  if (seed_dev) seed += (uint64_t)*seed_dev;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = lane / SEG, sl = lane % SEG;
  const bool active = FULL ? true : (sl * UPL < nO);
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const float inv_n = 1.f / (float)nO;
  const uint32_t thr = dropout_thr(drop_p);
  const int u0 = active ? sl * UPL : 0;
  float gk[UPL];
#pragma unroll
  for (int j = 0; j < UPL; ++j) gk[j] = has_ln ? bf2f(G[u0 + j]) : 1.f;
  float accG[UPL], accB[UPL], accb[ZPL];
#pragma unroll
  for (int j = 0; j < UPL; ++j) { accG[j] = 0.f; accB[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < ZPL; ++j) accb[j] = 0.f;
  const int ngroups = (Tp + R * RPW - 1) / (R * RPW);
  for (int grp = gwarp; grp < ngroups; grp += nwarps) {
    const int row0 = grp * R * RPW + seg;
    bf16x8 draw[R][UPL / 8], xraw[R][UPL / 8];
    uint2 wraw[R][UPL / 8];
    float mk[R], rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = min(row0 + r * RPW, Tp - 1);
      const size_t ro = (size_t)row * nO + u0;
      mk[r] = mask[row];
      rs[r] = has_ln ? rstd_in[row] : 1.f;
#pragma unroll
      for (int v = 0; v < UPL / 8; ++v) {
        draw[r][v] = *(const bf16x8*)(dY + ro + v * 8);
        if (has_ln) xraw[r][v] = *(const bf16x8*)(xhat + ro + v * 8);
        wraw[r][v] = *(const uint2*)(which + ro + v * 8);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * RPW;
      const bool row_ok = row < Tp;
      if (UNIFORM && !row_ok) break;
      if (UNIFORM && mk[r] == 0.0f) {           // pad row (warp-uniform): zeros
        if (active) {
          bf16x8 zero;
#pragma unroll
          for (int i = 0; i < 8; ++i) zero.v[i] = f2bf(0.f);
          bf16x8* zout = (bf16x8*)(dZ + (size_t)row * nO * NP + u0 * NP);
#pragma unroll
          for (int v = 0; v < ZPL / 8; ++v) zout[v] = zero;
        }
        continue;
      }
      // two rows per warp / idle lanes: a dead lane's d is forced to 0, which makes every one of its
      // accumulator contributions and its whole dZ slice exactly zero
      constexpr bool ALL_LIVE = UNIFORM && FULL;
      const bool live = UNIFORM ? active : (row_ok && active && mk[r] != 0.0f);
      const size_t ro = (size_t)(UNIFORM ? row : min(row, Tp - 1)) * nO + u0;
      float dn[UPL], xh[UPL];
      uint8_t wh[UPL];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int v = 0; v < UPL / 8; ++v) {
        const uint8_t* wb = (const uint8_t*)&wraw[r][v];
        float keep[8];
        if (drop_p > 0.f) dropout_scale8(seed, ro + v * 8, thr, inv_keep, keep);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = v * 8 + i;
          float d = (ALL_LIVE || live) ? bf2f(draw[r][v].v[i]) : 0.f;
          if (drop_p > 0.f) d *= keep[i];
          wh[j] = wb[i];
          if (has_ln) {
            xh[j] = (ALL_LIVE || live) ? bf2f(xraw[r][v].v[i]) : 0.f;
            accG[j] += d * xh[j];
            accB[j] += d;
            d *= gk[j];
            s1 += d; s2 += d * xh[j];
          } else {
            xh[j] = 0.f;
          }
          dn[j] = d;
        }
      }
      float rstd = 1.f;
      if (has_ln) {
        rstd = (ALL_LIVE || live) ? rs[r] : 0.f;
        s1 = seg_sum<SEG>(s1) * inv_n;
        s2 = seg_sum<SEG>(s2) * inv_n;
      }
      if (!active || !row_ok) continue;
      __align__(16) __nv_bfloat16 zo[ZPL];
#pragma unroll
      for (int j = 0; j < UPL; ++j) {
        const float dH = has_ln ? rstd * (dn[j] - s1 - xh[j] * s2) : dn[j];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const float val = (q == wh[j]) ? dH : 0.f;
          zo[j * NP + q] = f2bf(val);
          accb[j * NP + q] += val;
        }
      }
      bf16x8* zout = (bf16x8*)(dZ + (size_t)row * nO * NP + u0 * NP);
#pragma unroll
      for (int v = 0; v < ZPL / 8; ++v) zout[v] = *(const bf16x8*)(zo + v * 8);
    }
  }
  // two rows per warp: fold the upper segment's accumulators into the lower one first
  if (SEG == 16) {
#pragma unroll
    for (int j = 0; j < UPL; ++j) {
      accG[j] += __shfl_down_sync(0xffffffffu, accG[j], 16);
      accB[j] += __shfl_down_sync(0xffffffffu, accB[j], 16);
    }
#pragma unroll
    for (int j = 0; j < ZPL; ++j) accb[j] += __shfl_down_sync(0xffffffffu, accb[j], 16);
  }
  // combine the 8 warps of the block through shared memory, then one atomic per element
  extern __shared__ float sred[];                  // [8][nO*(NP+2)]
  const int per = nO * (NP + 2);
  float* mine = sred + (size_t)warp * per;
  if (active && seg == 0) {
#pragma unroll
    for (int j = 0; j < UPL; ++j) { mine[u0 + j] = accG[j]; mine[nO + u0 + j] = accB[j]; }
#pragma unroll
    for (int j = 0; j < ZPL; ++j) mine[2 * nO + u0 * NP + j] = accb[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < per; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sred[(size_t)w * per + i];
    if (t == 0.f) continue;
    if (i < nO) { if (has_ln) atomicAdd(dG + i, t); }
    else if (i < 2 * nO) { if (has_ln) atomicAdd(dbeta + (i - nO), t); }
    else atomicAdd(db + (i - 2 * nO), t);
  }
}

template <int NP, int UPL, int SEG, bool FULL>
static void launch_bwd_vec(const void* dY, const void* xhat, const float* rstd, const void* G, const uint8_t* which,
                           const float* mask, void* dZ, float* db, float* dG, float* dbeta, int Tp, int nO, float drop_p,
                           uint64_t seed, const int64_t* seed_dev, int has_ln, cudaStream_t s) {
  constexpr int R = rows_per_iter(UPL * 5);          // dY + xhat (2 B each) + which (1 B) per unit
  int blocks = (Tp + 31) / 32;                       // >= 4 rows per warp so the flush is amortised
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  const size_t smem = sizeof(float) * 8 * (size_t)nO * (NP + 2);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(maxout_ln_bwd_vec_kernel<NP, UPL, R, SEG, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  launch_k(maxout_ln_bwd_vec_kernel<NP, UPL, R, SEG, FULL>, blocks, 256, smem, s,
      (const __nv_bfloat16*)dY, (const __nv_bfloat16*)xhat, rstd, (const __nv_bfloat16*)G, which, mask,
      (__nv_bfloat16*)dZ, db, dG, dbeta, Tp, nO, drop_p, seed, seed_dev, has_ln);
}

bool try_launch_maxout_ln_bwd_vec(const void* dY, const void* xhat, const float* rstd, const void* G,
                                  const uint8_t* which, const float* mask, void* dZ, float* db, float* dG,
                                  float* dbeta, int Tp, int nO, int nP, float drop_p, uint64_t seed,
                                  const int64_t* seed_dev, int has_ln, cudaStream_t s) {
  int upl, seg;
  if (!vec_layout(nO, &upl, &seg)) return false;
#define SRB_TRY(NP_, UPL_, SEG_)                                                                                      \
  if (nP == NP_ && upl == UPL_ && seg == SEG_) {                                                                      \
    if (nO == UPL_ * SEG_)                                                                                            \
      launch_bwd_vec<NP_, UPL_, SEG_, true>(dY, xhat, rstd, G, which, mask, dZ, db, dG, dbeta, Tp, nO, drop_p, seed,  \
                                            seed_dev, has_ln, s);                                                     \
    else                                                                                                              \
      launch_bwd_vec<NP_, UPL_, SEG_, false>(dY, xhat, rstd, G, which, mask, dZ, db, dG, dbeta, Tp, nO, drop_p, seed, \
                                             seed_dev, has_ln, s);                                                    \
    return true;                                                                                                      \
  }
  SRB_TRY(3, 8, 32) SRB_TRY(3, 8, 16) SRB_TRY(3, 16, 32)
  SRB_TRY(2, 8, 32) SRB_TRY(2, 8, 16) SRB_TRY(2, 16, 32)
#undef SRB_TRY
  return false;
}

// ------------------------------------------------------------------------------------------
// HashEmbed backward over sorted ids
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int kSortChunk = 32;     // sorted positions per warp (one per lane in the prologue)

// keys: the raw (R, n_attr) attribute array; perm: (n_tables, R) rows grouped by id per table
// (device radix sort on truncated keys, or the host-side grouping shipped with the batch).
// Equal ids are adjacent; runs are split on the FULL 64-bit id.  A warp owns kSortChunk grouped
// positions: the dependent perm -> mask/key loads happen once, lane-parallel, in the prologue.
// One table row is only width/8 16-byte vectors, so the warp is split into G = 32/(width/8)
// lane groups, each walking its own contiguous slice of the chunk (its own runs, its own
// flushes) - with width 64 that is four positions in flight per step instead of one with 24
// idle lanes.  dY row loads are issued four positions deep per group.
template <typename PermT>
__global__ void __launch_bounds__(128) hash_embed_bwd_sorted_kernel(const int64_t* __restrict__ keys,
                                                                    const PermT* __restrict__ perm,
                                                                    const float* __restrict__ mask, HashEmbedTables t,
                                                                    const __nv_bfloat16* __restrict__ dY, int R) {
  pdl_prologue();
  const int a = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int p0 = chunk * kSortChunk;
  if (p0 >= R) return;
  const int n = min(R, p0 + kSortChunk) - p0;
  const int C = t.n_tables * t.width;
  const int nvec = t.width / 8;                      // 16-byte vectors per table row (<= 64)
  // lanes per position: nvec if it is a power of two below 32, else the whole warp (2 vectors/lane max)
  const int gsize = (nvec < 32 && (nvec & (nvec - 1)) == 0) ? nvec : 32;
  const int S = gsize;                               // positions per group = 32 / (32 / gsize)
  const int g0 = (lane / gsize) * S;                 // first chunk position of my group
  const int vl = lane % gsize;                       // my vector within the row
  const PermT* pm = perm + (size_t)a * R;
  const int col = t.column[a];
  int my_row = 0;
  float my_mask = 0.f;
  int64_t my_key = 0;
  if (lane < n) {
    my_row = (int)pm[p0 + lane];
    my_mask = mask[my_row];
    my_key = keys[(size_t)my_row * t.n_attr + col];
  }
  float acc[2][8];
#pragma unroll
  for (int v = 0; v < 2; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
  bool have = false;
  int64_t cur = 0;
  float* dE = t.grad[a];
  auto flush = [&]() {
    if (!have) return;
    uint32_t rows[4];
    hash_rows((uint64_t)cur, t.seed[a], t.n_rows[a], rows);
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int vec = vl + 32 * v;
      if (vec < nvec) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float* dst = dE + (size_t)rows[kk] * t.width + vec * 8;
          red_add_v4f(dst, acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
          red_add_v4f(dst + 4, acc[v][4], acc[v][5], acc[v][6], acc[v][7]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
      }
    }
  };
  constexpr int kBatch = 4;
  for (int j0 = 0; j0 < S; j0 += kBatch) {           // same trip count in every group: the shuffles are warp-wide
    int rowb[kBatch];
    float mb[kBatch];
    int64_t keyb[kBatch];
    bf16x8 g[kBatch][2];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int pos = g0 + j0 + j;                   // < 32 because S divides 32 and kBatch divides S or S < kBatch
      const int src = min(pos, kSortChunk - 1);
      rowb[j] = __shfl_sync(0xffffffffu, my_row, src);
      const float m = __shfl_sync(0xffffffffu, my_mask, src);
      keyb[j] = __shfl_sync(0xffffffffu, my_key, src);
      mb[j] = (j0 + j < S && pos < n) ? m : 0.f;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int vec = vl + 32 * v;
        if (vec < nvec && mb[j] != 0.f)
          g[j][v] = *(const bf16x8*)(dY + (size_t)rowb[j] * C + a * t.width + vec * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      if (mb[j] == 0.f) continue;                   // uniform within the lane group
      if (have && keyb[j] != cur) flush();
      cur = keyb[j]; have = true;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int vec = vl + 32 * v;
        if (vec < nvec) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[v][i] += bf2f(g[j][v].v[i]);
        }
      }
    }
  }
  flush();
}

void launch_hash_embed_bwd_sorted(const int64_t* keys, const void* perm, bool perm_is_i32, const float* mask,
                                  HashEmbedTables t, const void* dY, int R, cudaStream_t s) {
  if (R <= 0) return;
  dim3 grid((R + kSortChunk * 4 - 1) / (kSortChunk * 4), t.n_tables);
  if (perm_is_i32)
    launch_k(hash_embed_bwd_sorted_kernel<int32_t>, grid, 128, 0, s, keys, (const int32_t*)perm, mask, t,
                                                               (const __nv_bfloat16*)dY, R);
  else
    launch_k(hash_embed_bwd_sorted_kernel<int64_t>, grid, 128, 0, s, keys, (const int64_t*)perm, mask, t,
                                                               (const __nv_bfloat16*)dY, R);
}

// ------------------------------------------------------------------------------------------
// column sums of a tall (T, C) bf16 matrix, accumulated into fp32: bias gradients.
// (torch's sum(dim=0) on tall-skinny inputs costs 20-40 us plus an fp32 copy of the input.)
// A block covers a contiguous row range; thread = (row sub-group, 16-byte column vector).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, float* __restrict__ out,
                                                          int T, int C, int ld, int rows_per_block, int n_valid) {
  pdl_prologue();
  extern __shared__ float cs_red[];                 // [rsubs][C]
  const int c8 = C / 8;
  const int rsubs = blockDim.x / c8;
  const int vec = threadIdx.x % c8, rsub = threadIdx.x / c8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(T, r0 + rows_per_block);
  if (rsub < rsubs) {
    int r = r0 + rsub;
    for (; r + 3 * rsubs < r1; r += 4 * rsubs) {     // four independent loads in flight
      bf16x8 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *(const bf16x8*)(X + (size_t)(r + u * rsubs) * ld + vec * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += bf2f(v[u].v[i]);
    }
    for (; r < r1; r += rsubs) {
      bf16x8 v = *(const bf16x8*)(X + (size_t)r * ld + vec * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += bf2f(v.v[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) cs_red[(size_t)rsub * C + vec * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int j = 0; j < rsubs; ++j) t += cs_red[(size_t)j * C + c];
    if (t != 0.f && c < n_valid) atomicAdd(out + c, t);      // columns >= n_valid are pitch padding
  }
}

bool try_launch_colsum_bf16(const void* X, float* out, int T, int C, int ld, int n_valid, cudaStream_t s) {
  if (C % 8 != 0 || C / 8 > 256 || C <= 0 || T <= 0 || ld % 8 != 0) return false;
  const int c8 = C / 8, rsubs = 256 / c8;
  int blocks = 148 * 4;
  int rows_per_block = (T + blocks - 1) / blocks;
  if (rows_per_block < 4 * rsubs) rows_per_block = 4 * rsubs;
  blocks = (T + rows_per_block - 1) / rows_per_block;
  const size_t smem = sizeof(float) * (size_t)rsubs * C;     // <= 256 * 8 * 4 = 8 KB
  launch_k(colsum_bf16_kernel, blocks, 256, smem, s, (const __nv_bfloat16*)X, out, T, C, ld, rows_per_block,
                                               n_valid > 0 ? n_valid : C);
  return true;
}

// ------------------------------------------------------------------------------------------
// dst (bf16) = src (fp32); src = 0.  One pass: the fp32 scatter target of the transition backward
// (dYf) is converted for the tcgen05 GEMMs that consume it AND left clear for the next step, so
// the step needs neither a memset of the accumulator nor a separate dtype copy.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) f32_to_bf16_zero_kernel(float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                               size_t n8) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float4* p = (float4*)(src + i * 8);
    const float4 a = p[0], b = p[1];
    bf16x8 o;
    o.v[0] = f2bf(a.x); o.v[1] = f2bf(a.y); o.v[2] = f2bf(a.z); o.v[3] = f2bf(a.w);
    o.v[4] = f2bf(b.x); o.v[5] = f2bf(b.y); o.v[6] = f2bf(b.z); o.v[7] = f2bf(b.w);
    *(bf16x8*)(dst + i * 8) = o;
    p[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    p[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

void launch_f32_to_bf16_zero(float* src, void* dst, size_t n, cudaStream_t s) {
  const size_t n8 = n / 8;
  if (n8 == 0) return;
  size_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(f32_to_bf16_zero_kernel, (unsigned)blocks, 256, 0, s, src, (__nv_bfloat16*)dst, n8);
}

// ------------------------------------------------------------------------------------------
// One launch initialises a whole record arena: the first n_zero 32-bit words to 0, the next n_ones
// words to 0xFFFFFFFF (int32 -1: "missing feature" / "no action").  Replaces the 3-7 library fill
// kernels per transition head and step (hid, d_scores, which, loss, feats, history, n_steps).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) arena_init_kernel(uint32_t* __restrict__ p, size_t n_zero, size_t n_ones) {
  pdl_prologue();
  const size_t n = n_zero + n_ones;
  const size_t n4 = n / 4;                                   // the arena is 16-byte aligned and padded
  const size_t z4 = n_zero / 4;                              // n_zero is a multiple of 4 words
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t v = i < z4 ? 0u : 0xFFFFFFFFu;
    ((uint4*)p)[i] = make_uint4(v, v, v, v);
  }
}

void launch_arena_init(void* p, size_t n_zero_words, size_t n_ones_words, cudaStream_t s) {
  const size_t n4 = (n_zero_words + n_ones_words) / 4;
  if (n4 == 0) return;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(arena_init_kernel, (unsigned)blocks, 256, 0, s, (uint32_t*)p, n_zero_words, n_ones_words);
}

// the device-side dropout stream position: one thread of ours instead of a library elementwise kernel
__global__ void bump_i64_kernel(int64_t* p, int64_t by) {
  pdl_prologue();
  *p += by;
}
void launch_bump_i64(int64_t* p, int64_t by, cudaStream_t s) { launch_k(bump_i64_kernel, 1, 1, 0, s, p, by); }

}  // namespace srb
