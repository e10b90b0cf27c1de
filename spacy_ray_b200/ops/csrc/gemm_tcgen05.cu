// Persistent warp-specialised tcgen05 GEMM family for the tok2vec / parser hot path.
//
//   D[M, N] = sum over "shifts" s and k of  A[m + a_row_shift[s], a_col_off[s] + k] * B[n + b_row_off[s], b_col_off[s] + k]
//
// * bf16 operands staged in shared memory by TMA (SWIZZLE_128B), fp32 accumulators in
//   TMEM (two stages, so the epilogue of tile i overlaps the MMAs of tile i+1),
//   tcgen05.mma issued by one elected thread, tcgen05.ld epilogue.
// * The shift table is how expand_window (seq2col, K4 in SURVEY.md 2.7) disappears: the
//   window GEMM loads the SAME activation array three times at row offsets -1/0/+1
//   (the batch layout keeps a zero row between docs, so TMA out-of-bounds/zero rows
//   give the doc-edge padding for free) instead of materialising (T, 3*width).
// * Fused epilogues: +bias -> bf16 store (Linear / precompute), +bias -> maxout over
//   3 pieces -> bf16 + argmax byte (K2), + residual*mask add (window dX), fp32
//   split-K reduction straight into the flat gradient buffer (dW; both operands
//   MN-major so no transposes are materialised).
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..9 = epilogue (TMEM lane quadrant = warp_id % 4, column half = (warp_id - 2) / 4).
#include <cstdlib>
#include "common.cuh"
#include "gate.cuh"
#include "gemm_launch.h"
#include "launch.h"
#include "gemm_tcgen05.cuh"

namespace srb {

using namespace sm100;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumThreads = 320;          // TMA warp + MMA warp + 8 epilogue warps
constexpr int kSmemBudget = 200 * 1024;
constexpr int kMaxBiasN = 4096;           // bias staged in smem as fp32 (16 KB)

// PAIR (cta_group::2): a CTA stores only its half of the B tile, so a stage is smaller and the
// ring is deeper - more bytes in flight per SM for the same shared memory.
// HALO (window GEMMs, shifts -1/0/+1): the A tile is loaded ONCE per k-block with one extra row
// above and below (130 rows x 128 B) and the three shifted operands are read out of it through
// descriptors that start 0, 1 or 2 rows (128 B each) into the tile.  The 128 B swizzle is a function
// of the absolute shared-memory address bits, so a row-shifted start address needs no further
// correction (measured: base_offset must stay 0; setting it to the row phase gives wrong data).  That removes two
// of the three A loads of these L2->SM-bound kernels; a stage then carries the B tiles of all
// three shifts.
// BRES (window forward, pair MMA, K <= 256): the B operand - the weights - stays RESIDENT in shared
// memory for the whole kernel.  With a grid of clusters that is a multiple of the number of N tiles,
// the persistent schedule (item = cluster + i * clusters, tile = m_group * n_tiles + n_tile) gives
// every cluster ONE fixed N tile, i.e. one 192-row slice of W: 4 k-blocks x 3 shifts x 12 KB = 144 KB
// per CTA, loaded once.  The ring then carries only the 17 KB activation tiles.  Before, every tile
// re-fetched its 144 KB of weights from L2 (119 MB per layer against 53 MB of activations) and the
// kernel sat at the L2 -> SM cap with the tensor pipe 68 % busy.
template <int BLOCK_N, bool PAIR = false, bool HALO = false, bool BRES = false>
struct Cfg {
  static constexpr int kTileB = (PAIR ? BLOCK_N / 2 : BLOCK_N) * BK * 2;
  static constexpr int kBoxA = (HALO ? kHaloRows : BM) * BK * 2;             // bytes the A load delivers
  static constexpr int kStageA = HALO ? 17 * 1024 : BM * BK * 2;             // 1024-aligned
  static constexpr int kStageB = BRES ? 0 : (HALO ? 3 : 1) * kTileB;
  static constexpr int kResKb = 4;                                           // k-blocks of 64 held resident (K <= 256)
  static constexpr int kResB = BRES ? kResKb * 3 * kTileB : 0;
  static constexpr int kTxBytes = kBoxA + kStageB;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kBiasBytes = BRES ? 8192 : kMaxBiasN * 4;             // bias (+ LayerNorm gain/shift) as fp32
  static constexpr int kStagesRaw = BRES ? 4 : kSmemBudget / kStage;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStage + kResB + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- EPI_MAXOUT3_LN, second phase (see gemm_launch.h and the epilogue below) ------------------
// What an epilogue thread carries from the first phase of a tile: its 32 maxout outputs of one row
// (bf16 pairs), their argmax pieces (2 bits each) and where the tile sits.
struct LnTile {
  uint32_t hp[16];
  uint32_t wb[2];
  int m0, ntile;
};

__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// What the second phase needs from memory, requested in one go (ONE L2 round trip per tile).  Nothing is
// consumed here: the raw 16-byte entries stay in registers until ln_finish.
struct LnLoads {
  float mk;
  uint4 xr[4];
  float4 st[4];                 // the row's first four tile entries (rows with more tiles: the rest is read late)
};

__device__ __forceinline__ float4 ld_shared_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

template <int BLOCK_N>
__device__ __forceinline__ void ln_prefetch(const GemmParams& p, const LnTile& t, LnLoads& ld, int M, int n_tiles, int q,
                                            int half, int lane) {
  const LnArgs& L = p.ln;
  const int row = t.m0 + q * 32 + lane;
  if (row >= M) return;
  const int unit0 = (t.ntile * BLOCK_N + half * (BLOCK_N / 2)) / 3;
  ld.mk = L.mask[row];
  if (L.xres) {
    const uint4* xp = (const uint4*)(L.xres + (size_t)row * L.ld_res + unit0);
#pragma unroll
    for (int g = 0; g < 4; ++g) ld.xr[g] = xp[g];
  }
  const float4* sp = L.stats + (size_t)row * n_tiles;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < n_tiles) ld.st[j] = ld_cg_f4(sp + j);
}

template <int BLOCK_N>
__device__ __forceinline__ void ln_finish(const GemmParams& p, const LnTile& t, const LnLoads& ld, const float* bias_s,
                                          int M, int n_tiles, int q, int half, int lane, uint32_t tag) {
  constexpr int HN = BLOCK_N / 2;
  const LnArgs& L = p.ln;
  const int nO = p.N / 3;
  const int unit0 = (t.ntile * BLOCK_N + half * HN) / 3;          // my 32 units of the row
  const int row = t.m0 + q * 32 + lane;
  if (row >= M) return;                                            // nobody waits for rows past the end
  __nv_bfloat16* out = (__nv_bfloat16*)p.out;
  const size_t yo = (size_t)row * p.ldo + unit0;
  const size_t ro = (size_t)row * nO + unit0;                      // xhat / which / dropout index: dense (rows, nO)
  const float mk = ld.mk;
  const uint4* xr = ld.xr;
  const float4* sp = L.stats + (size_t)row * n_tiles;
  float t1 = 0.f, t2 = 0.f;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < n_tiles) { ok = ok && (__float_as_uint(ld.st[j].z) == tag); t1 += ld.st[j].x; t2 += ld.st[j].y; }
  for (int j = 4; j < n_tiles; ++j) {
    const float4 e = ld_cg_f4(sp + j);
    ok = ok && (__float_as_uint(e.z) == tag);
    t1 += e.x; t2 += e.y;
  }
  if (!ok) {                                                       // rare: some tile of the row was not published yet
    const uint64_t t0 = globaltimer_ns();
    while (true) {
      ok = true;
      t1 = 0.f; t2 = 0.f;
      for (int j = 0; j < n_tiles; ++j) {
        const float4 e = ld_cg_f4(sp + j);
        ok = ok && (__float_as_uint(e.z) == tag);
        t1 += e.x; t2 += e.y;
      }
      if (ok) break;
      if (globaltimer_ns() - t0 > 2000000000ull) { atomicExch(L.seq + 2, 1u); break; }
      __nanosleep(64);
    }
  }
  if (mk == 0.0f) {                                                // pad row: zeros (the next window GEMM relies on them)
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int g = 0; g < 4; ++g) { *(uint4*)(out + yo + g * 8) = z; *(uint4*)(L.xhat + ro + g * 8) = z; }
    *(uint4*)(p.which + ro) = z; *(uint4*)(p.which + ro + 16) = z;
    if (t.ntile == 0 && half == 0) L.rstd[row] = 0.f;
    return;
  }
  const float inv_n = 1.f / (float)nO;
  const float mu = t1 * inv_n;
  const float rstd = rsqrtf(fmaxf(t2 * inv_n - mu * mu, 0.f) + 1e-8f);
  uint64_t seed = L.seed;
  if (L.seed_dev) seed += (uint64_t)*L.seed_dev;
  const float inv_keep = L.drop_p > 0.f ? 1.0f / (1.0f - L.drop_p) : 1.0f;
  const uint32_t thr = dropout_thr(L.drop_p);
  const float* g4 = bias_s + p.N + unit0;                          // LayerNorm gain / shift of my units (smem, fp32)
  const float* b4 = g4 + nO;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float keep[8];
    if (L.drop_p > 0.f) dropout_scale8(seed, ro + g * 8, thr, inv_keep, keep);
    const float4 ga = ld_shared_f4(g4 + 8 * g), gb = ld_shared_f4(g4 + 8 * g + 4);
    const float4 ba = ld_shared_f4(b4 + 8 * g), bb = ld_shared_f4(b4 + 8 * g + 4);
    const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    const uint32_t xw[4] = {xr[g].x, xr[g].y, xr[g].z, xr[g].w};
    float yv[8], xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = g * 8 + i;
      const uint32_t pr = t.hp[j >> 1];
      const float hf = __uint_as_float((j & 1) ? (pr & 0xFFFF0000u) : (pr << 16));
      const float xh = (hf - mu) * rstd;
      float nv = fmaf(xh, gg[i], bv[i]);
      if (L.drop_p > 0.f) nv *= keep[i];
      if (L.xres) nv += __uint_as_float((i & 1) ? (xw[i >> 1] & 0xFFFF0000u) : (xw[i >> 1] << 16));
      yv[i] = nv; xv[i] = xh;
    }
    *(uint4*)(out + yo + g * 8) = make_uint4(pack_bf16x2(yv[0], yv[1]), pack_bf16x2(yv[2], yv[3]),
                                             pack_bf16x2(yv[4], yv[5]), pack_bf16x2(yv[6], yv[7]));
    *(uint4*)(L.xhat + ro + g * 8) = make_uint4(pack_bf16x2(xv[0], xv[1]), pack_bf16x2(xv[2], xv[3]),
                                                pack_bf16x2(xv[4], xv[5]), pack_bf16x2(xv[6], xv[7]));
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    uint32_t w4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t f = t.wb[cc] >> (8 * k);                      // four 2-bit fields -> four bytes
      w4[k] = (f & 3u) | ((f & 12u) << 6) | ((f & 48u) << 12) | ((f & 192u) << 18);
    }
    *(uint4*)(p.which + ro + cc * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
  if (t.ntile == 0 && half == 0) L.rstd[row] = rstd;
}

// CL = CTAs per cluster (1 or 2).  With CL == 2 the two CTAs of a cluster work on vertically
// adjacent M tiles of the SAME N tile / K range in lockstep and share the B (weight) operand:
// each CTA TMA-loads half of the B tile and multicasts it into both CTAs' shared memory, so B
// crosses L2 -> SM once per cluster instead of once per CTA (the single-CTA kernels measured
// ~55-60 % tensor-pipe activity with L2 -> smem traffic as the limiter).
//
// PAIR = the two CTAs additionally issue ONE tcgen05.mma.cta_group::2 per k-step (M = 256: each
// CTA contributes its 128 A rows and HALF of the B tile from its own shared memory; each CTA's
// TMEM receives its own 128 accumulator rows).  Per SM and per FLOP that halves the B bytes
// stored AND fetched, which is what the L2 -> smem-bound single-CTA kernels were missing.  Both
// CTAs run a TMA producer (completion bytes are signalled on the leader's barriers), only the
// leader (cluster rank 0) runs the MMA issuer, both run their own epilogue.
template <int BLOCK_N, int MODE, int EPI, int CL, bool PAIR, bool HALO = false, bool BRES = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  static_assert(!PAIR || CL == 2, "the pair MMA needs a 2-CTA cluster");
  static_assert(!HALO || (MODE != MODE_MNMN && (PAIR || CL == 1)), "halo: K-major A, single CTA or pair MMA");
  static_assert(!BRES || (HALO && PAIR && MODE == MODE_KK), "resident B: window forward with the pair MMA");
  using C = Cfg<BLOCK_N, PAIR, HALO, BRES>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + C::kStages * C::kStageA;
  unsigned char* sBres = smem + C::kStages * C::kStage;       // BRES: the cluster's slice of the weights, all k-blocks
  uint64_t* full_bar = (uint64_t*)(smem + C::kStages * C::kStage + C::kResB);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* bres_bar = tmem_empty + 2;
  uint32_t* tmem_ptr = (uint32_t*)(bres_bar + 1);
  float* bias_s = (float*)(smem + C::kStages * C::kStage + C::kResB + 256);

  pdl_trigger();          // the next kernel of the stream may be scheduled behind this one's CTAs (launch.h)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;    // a partial last N tile is masked in the epilogue (TMA zero-fills B)
  const int kpb = (p.K + BK - 1) / BK;                   // k-blocks per shift
  const int total_kb = (MODE != MODE_MNMN && !HALO ? p.n_shifts : 1) * kpb;
  const int splits = p.splits > 0 ? p.splits : 1;
  const int m_groups = (m_tiles + CL - 1) / CL;              // a cluster takes CL vertically adjacent M tiles
  const int total_items = m_groups * n_tiles * splits;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  const int first_item = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const int item_stride = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    // C2: the all-gather's consumer side.  The owners of this layer's weights pushed them into my
    // weight buffer from inside their exchange kernels; wait for their "published" flags here, while
    // warp 1 initialises barriers and allocates TMEM, instead of at the end of the previous step.
    gate_wait_warp(p.gate);
  }
  if (warp == 1 && lane == 0) {
    // PAIR: one MMA issuer (the leader) frees a stage in both CTAs; its tmem_empty barrier
    // collects the epilogue warps of BOTH CTAs
    for (int i = 0; i < C::kStages; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, PAIR ? 1 : CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, PAIR ? 16 : 8); }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { if (PAIR) tmem_alloc_pair<C::kTmemCols>(tmem_ptr); else tmem_alloc<C::kTmemCols>(tmem_ptr); }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();      // peers' barriers are initialised too
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // everything above touched no global memory (descriptors are kernel parameters; gate flags are
  // written by peers, not by the stream predecessor): with programmatic dependent launch it ran
  // under the tail of the previous kernel.  From here on its outputs are read and buffers it may
  // still be reading are written.
  pdl_wait();
  const int M = p.m_dev ? min(p.M, *p.m_dev) : p.M;
  const uint32_t ln_tag = EPI == EPI_MAXOUT3_LN ? *(volatile const uint32_t*)p.ln.seq : 0u;
  // gated launch: peers' generic-proxy stores (the freshly published bucket) -> my TMA loads
  if (p.gate.flags != nullptr && warp == 0) fence_proxy_async_global();
  if ((p.bias || EPI == EPI_MAXOUT3_LN) && warp >= 2) {      // bias (and LayerNorm gain / shift) -> smem (fp32) once per CTA
    if (p.bias)
      for (int i = threadIdx.x - 64; i < p.N; i += kNumThreads - 64) bias_s[i] = bf2f(p.bias[i]);
    if (EPI == EPI_MAXOUT3_LN) {
      const int nO = p.N / 3;
      for (int i = threadIdx.x - 64; i < nO; i += kNumThreads - 64) {
        bias_s[p.N + i] = bf2f(p.ln.G[i]);
        bias_s[p.N + nO + i] = bf2f(p.ln.beta[i]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kNumThreads - 64) : "memory");
  }

  if (warp == 0) {
    // ================================ TMA producer =====================================
    if (elect_one()) {
      int stage = 0, phase = 0;
      if (BRES && first_item < total_items) {
        // my N tile never changes (grid of clusters = multiple of n_tiles): fetch its weights ONCE
        const int n0r = (first_item % n_tiles) * BLOCK_N + (int)cta_rank * (BLOCK_N / 2);
        const uint32_t rb = map_to_cta(bres_bar, 0);
        if (cta_rank == 0) mbar_arrive_expect_tx(bres_bar, 2 * total_kb * 3 * C::kTileB);
        for (int kb = 0; kb < total_kb; ++kb)
#pragma unroll
          for (int s = 0; s < 3; ++s)
            tma_load_2d_2sm(sBres + (kb * 3 + s) * C::kTileB, &tmB, rb, p.b_col_off[s] + kb * BK, n0r + p.b_row_off[s]);
      }
      for (int w = first_item; w < total_items; w += item_stride) {
        const int tile = w / splits, split = w - tile * splits;
        const int mg = tile / n_tiles;
        const int m0 = (mg * CL + (int)cta_rank) * BM, n0 = (tile % n_tiles) * BLOCK_N;
        if (mg * CL * BM >= M) continue;                       // decided per cluster, not per CTA
        const int kb0 = (int)((long long)total_kb * split / splits), kb1 = (int)((long long)total_kb * (split + 1) / splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          unsigned char* a_dst = sA + stage * C::kStageA;
          unsigned char* b_dst = sB + stage * C::kStageB;
          if (HALO) {
            // one A load (130 rows starting at m0 - 1; rows outside the array are zero-filled) and the
            // B tiles of the three shifts
            const uint32_t fb = PAIR ? map_to_cta(full_bar + stage, 0) : 0u;
            if (!PAIR || cta_rank == 0) mbar_arrive_expect_tx(full_bar + stage, (PAIR ? 2 : 1) * C::kTxBytes);
            if (PAIR) tma_load_2d_2sm(a_dst, &tmA, fb, p.a_col_off[0] + kb * BK, m0 - 1);
            else tma_load_2d(a_dst, &tmA, full_bar + stage, p.a_col_off[0] + kb * BK, m0 - 1);
            if (!BRES)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              unsigned char* bs = b_dst + s * C::kTileB;
              if (MODE == MODE_KK) {
                const int c = p.b_col_off[s] + kb * BK, r = n0 + p.b_row_off[s] + (PAIR ? (int)cta_rank * (BLOCK_N / 2) : 0);
                if (PAIR) tma_load_2d_2sm(bs, &tmB, fb, c, r);
                else tma_load_2d(bs, &tmB, full_bar + stage, c, r);
              } else {                               // MODE_KMN: 64-column atoms of W as stored
                constexpr int kAtoms = BLOCK_N / 64 / (PAIR ? 2 : 1);
                const int krow = p.b_row_off[s] + kb * BK, c0 = p.b_col_off[s] + n0;
#pragma unroll
                for (int jj = 0; jj < kAtoms; ++jj) {
                  const int j = (PAIR ? (int)cta_rank * kAtoms : 0) + jj;
                  if (PAIR) tma_load_2d_2sm(bs + jj * (BK * 128), &tmB, fb, c0 + j * 64, krow);
                  else tma_load_2d(bs + jj * (BK * 128), &tmB, full_bar + stage, c0 + j * 64, krow);
                }
              }
            }
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (PAIR) {
            // both CTAs' bytes are counted on the LEADER's full barrier
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * C::kStage);
            const uint32_t fb = map_to_cta(full_bar + stage, 0);
            int s = 0, kk = kb;
            if (MODE != MODE_MNMN) { s = kb / kpb; kk = kb - s * kpb; }
            if (MODE == MODE_KK) {
              constexpr int kHalfRows = BLOCK_N / 2;
              tma_load_2d_2sm(a_dst, &tmA, fb, p.a_col_off[s] + kk * BK, m0 + p.a_row_shift[s]);
              tma_load_2d_2sm(b_dst, &tmB, fb, p.b_col_off[s] + kk * BK, n0 + p.b_row_off[s] + (int)cta_rank * kHalfRows);
            } else {
              constexpr int kAtoms = BLOCK_N / 64 / 2;            // my 64-column atoms of the B tile
              int c0, krow;
              if (MODE == MODE_KMN) {
                tma_load_2d_2sm(a_dst, &tmA, fb, p.a_col_off[s] + kk * BK, m0 + p.a_row_shift[s]);
                krow = p.b_row_off[s] + kk * BK; c0 = p.b_col_off[s] + n0;
              } else {
                const int t0 = kb * BK;
                int tshift = 0;
                c0 = n0;
                if (p.win_w > 0) { const int sw = n0 / p.win_w; c0 = n0 - sw * p.win_w; tshift = sw - 1; }
                krow = t0 + tshift;
#pragma unroll
                for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(a_dst + j * (BK * 128), &tmA, fb, m0 + j * 64, t0);
              }
#pragma unroll
              for (int jj = 0; jj < kAtoms; ++jj)
                tma_load_2d_2sm(b_dst + jj * (BK * 128), &tmB, fb, c0 + ((int)cta_rank * kAtoms + jj) * 64, krow);
            }
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar + stage, C::kStage);
          if (MODE == MODE_KK) {
            const int s = kb / kpb, kk = kb - s * kpb;
            tma_load_2d(a_dst, &tmA, full_bar + stage, p.a_col_off[s] + kk * BK, m0 + p.a_row_shift[s]);
            if (CL == 1) {
              tma_load_2d(b_dst, &tmB, full_bar + stage, p.b_col_off[s] + kk * BK, n0 + p.b_row_off[s]);
            } else {       // my half of the B rows, delivered to every CTA of the cluster
              constexpr int kHalfRows = BLOCK_N / CL;
              tma_load_2d_mcast(b_dst + cta_rank * (kHalfRows * BK * 2), &tmB, full_bar + stage,
                                p.b_col_off[s] + kk * BK, n0 + p.b_row_off[s] + (int)cta_rank * kHalfRows, kMask);
            }
          } else if (MODE == MODE_KMN) {
            // A as in MODE_KK; B = the weight matrix as stored, (K, N) row-major: one 64-column
            // (N) atom per TMA box, BK rows of K each
            const int s = kb / kpb, kk = kb - s * kpb;
            tma_load_2d(a_dst, &tmA, full_bar + stage, p.a_col_off[s] + kk * BK, m0 + p.a_row_shift[s]);
            const int krow = p.b_row_off[s] + kk * BK, c0 = p.b_col_off[s] + n0;
            if (CL == 1) {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_2d(b_dst + j * (BK * 128), &tmB, full_bar + stage, c0 + j * 64, krow);
            } else {
              constexpr int kAtoms = BLOCK_N / 64 / CL;
#pragma unroll
              for (int jj = 0; jj < kAtoms; ++jj) {
                const int j = (int)cta_rank * kAtoms + jj;
                tma_load_2d_mcast(b_dst + j * (BK * 128), &tmB, full_bar + stage, c0 + j * 64, krow, kMask);
              }
            }
          } else {
            const int t0 = kb * BK;
            int c0 = n0, tshift = 0;
            if (p.win_w > 0) { const int s = n0 / p.win_w; c0 = n0 - s * p.win_w; tshift = s - 1; }
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * (BK * 128), &tmA, full_bar + stage, m0 + j * 64, t0);
            if (CL == 1) {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_2d(b_dst + j * (BK * 128), &tmB, full_bar + stage, c0 + j * 64, t0 + tshift);
            } else {       // my share of the 64-column atoms, multicast to the cluster
              constexpr int kAtoms = BLOCK_N / 64 / CL;
#pragma unroll
              for (int jj = 0; jj < kAtoms; ++jj) {
                const int j = (int)cta_rank * kAtoms + jj;
                tma_load_2d_mcast(b_dst + j * (BK * 128), &tmB, full_bar + stage, c0 + j * 64, t0 + tshift, kMask);
              }
            }
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =======================================
    if ((!PAIR || cta_rank == 0) && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BLOCK_N, MODE == MODE_MNMN, MODE != MODE_KK);
      int stage = 0, phase = 0, it = 0;
      if (BRES && first_item < total_items) { mbar_wait(bres_bar, 0); tc_fence_after(); }
      for (int w = first_item; w < total_items; w += item_stride) {
        const int tile = w / splits, split = w - tile * splits;
        const int mg = tile / n_tiles;
        if (mg * CL * BM >= M) continue;
        const int kb0 = (int)((long long)total_kb * split / splits), kb1 = (int)((long long)total_kb * (split + 1) / splits);
        const int acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::kStageA);
          const uint32_t b_addr = smem_u32(sB + stage * C::kStageB);
          if (HALO) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint32_t r = (uint32_t)(p.a_row_shift[s] + 1);           // 0, 1 or 2 rows into the halo tile
              const uint32_t bs = BRES ? smem_u32(sBres) + ((kb - kb0) * 3 + s) * C::kTileB : b_addr + s * C::kTileB;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t adesc = make_smem_desc(a_addr + r * 128 + k * 32, 0, 1024);
                const uint64_t bdesc = MODE == MODE_KK ? make_smem_desc(bs + k * 32, 0, 1024)
                                                       : make_smem_desc(bs + k * 16 * 128, BK * 128, 1024);
                const uint32_t accum = (kb > kb0 || s > 0 || k > 0) ? 1u : 0u;
                if (PAIR) umma_bf16_pair(d_tmem, adesc, bdesc, idesc, accum);
                else umma_bf16(d_tmem, adesc, bdesc, idesc, accum);
              }
            }
          } else {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t adesc, bdesc;
            if (MODE == MODE_KK) {
              adesc = make_smem_desc(a_addr + k * 32, 0, 1024);
              bdesc = make_smem_desc(b_addr + k * 32, 0, 1024);
            } else if (MODE == MODE_KMN) {
              adesc = make_smem_desc(a_addr + k * 32, 0, 1024);
              bdesc = make_smem_desc(b_addr + k * 16 * 128, BK * 128, 1024);
            } else {
              adesc = make_smem_desc(a_addr + k * 16 * 128, BK * 128, 1024);
              bdesc = make_smem_desc(b_addr + k * 16 * 128, BK * 128, 1024);
            }
            if (PAIR) umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          }
          if (PAIR) umma_commit_pair(empty_bar + stage, kMask);
          else if (CL == 1) umma_commit(empty_bar + stage);
          else umma_commit_mcast(empty_bar + stage, kMask);     // frees the stage in every CTA of the cluster
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair(tmem_full + acc, kMask);    // both CTAs' epilogues own rows of this tile
        else umma_commit(tmem_full + acc);
        ++it;
      }
    }
  } else {
    // ================================ epilogue =========================================
    // 8 warps: warp -> TMEM lane quadrant (warp % 4) and column half ((warp - 2) / 4).  The MMA
    // warp was measured spinning on tmem_empty (4 warps, per-element bias LDGs, one wait per
    // tcgen05.ld): the epilogue, not the tensor pipe, paced the kernel.
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;             // which half of the tile's columns
    constexpr int HN = BLOCK_N / 2;
    int it = 0;
    LnTile ln_prev;                                  // EPI_MAXOUT3_LN: the tile whose second phase is still due
    LnLoads ln_ld;
    bool ln_have = false;
    for (int w = first_item; w < total_items; w += item_stride) {
      const int tile = w / splits;
      const int mg = tile / n_tiles;
      const int m0 = (mg * CL + (int)cta_rank) * BM, n0 = (tile % n_tiles) * BLOCK_N + half * HN;
      if (mg * CL * BM >= M) continue;
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t t_base = tmem_base + acc * BLOCK_N + half * HN + ((uint32_t)(q * 32) << 16);
      if (EPI == EPI_STORE) {
        __nv_bfloat16* out = (__nv_bfloat16*)p.out;
        const float rs = (p.add_src && p.row_scale && row_ok) ? p.row_scale[row] : 1.f;
#pragma unroll 1
        for (int c = 0; c < HN; c += 32) {
          float v[32];
          tmem_ld_x16_nowait(t_base + c, v);
          if (c + 16 < HN) tmem_ld_x16_nowait(t_base + c + 16, v + 16);
          tmem_ld_wait_regs<32>(v);
          if (row_ok) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (c + hh * 16 >= HN) break;
              float* vv = v + hh * 16;
              const int cc = c + hh * 16;
              if (n0 + cc >= p.N) break;                 // partial last N tile (N % 16 == 0): nothing to store
              if (p.bias) {                               // explicit ld.shared (the pointer is generic to the compiler)
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 bq = ld_shared_f4(bias_s + n0 + cc + i);
                  vv[i] += bq.x; vv[i + 1] += bq.y; vv[i + 2] += bq.z; vv[i + 3] += bq.w;
                }
              }
              if (p.add_src) {
                const bf16x8* src = (const bf16x8*)(p.add_src + (size_t)row * p.ld_add + n0 + cc);
                bf16x8 s0 = src[0], s1 = src[1];
#pragma unroll
                for (int i = 0; i < 8; ++i) { vv[i] += rs * bf2f(s0.v[i]); vv[8 + i] += rs * bf2f(s1.v[i]); }
              }
              bf16x8 o0, o1;
#pragma unroll
              for (int i = 0; i < 8; ++i) { o0.v[i] = f2bf(vv[i]); o1.v[i] = f2bf(vv[8 + i]); }
              bf16x8* dst = (bf16x8*)(out + (size_t)row * p.ldo + n0 + cc);
              dst[0] = o0; dst[1] = o1;
            }
          }
        }
      } else if (EPI == EPI_MAXOUT3) {
        __nv_bfloat16* out = (__nv_bfloat16*)p.out;
#pragma unroll 1
        for (int c = 0; c < HN; c += 48) {
          float v[48];
          tmem_ld_x16_nowait(t_base + c, v);
          tmem_ld_x16_nowait(t_base + c + 16, v + 16);
          tmem_ld_x16_nowait(t_base + c + 32, v + 32);
          tmem_ld_wait_regs<48>(v);
          if (row_ok) {
            bf16x8 h0, h1;
            __align__(16) uint8_t wh[16];
            if (p.bias) {
              // 48 consecutive bias values as 12 x ld.shared.v4 (the compiler only sees a generic pointer here:
              // the scalar form compiled to 48 generic LD.E per chunk)
#pragma unroll
              for (int k = 0; k < 12; ++k) {
                const float4 bq = ld_shared_f4(bias_s + n0 + c + 4 * k);
                v[4 * k] += bq.x; v[4 * k + 1] += bq.y; v[4 * k + 2] += bq.z; v[4 * k + 3] += bq.w;
              }
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float a = v[3 * u], b = v[3 * u + 1], d = v[3 * u + 2];
              float best = a; int bi = 0;
              if (b > best) { best = b; bi = 1; }
              if (d > best) { best = d; bi = 2; }
              if (u < 8) h0.v[u] = f2bf(best); else h1.v[u - 8] = f2bf(best);
              wh[u] = (uint8_t)bi;
            }
            const size_t uo = (size_t)row * p.ldo + (n0 + c) / 3;
            bf16x8* dst = (bf16x8*)(out + uo);
            dst[0] = h0; dst[1] = h1;
            *(uint4*)(p.which + uo) = *(const uint4*)wh;
          }
        }
      } else if (EPI == EPI_MAXOUT3_LN) {
        // ---- maxout + LayerNorm + dropout + residual + mask, the row never leaves the SM (gemm_launch.h) ----
        // Software-pipelined by one tile: phase A of tile i (accumulator -> maxout -> partial sums published)
        // runs BEFORE phase B of tile i-1 (wait for the other tiles' sums, normalise, store), so the other
        // clusters have a whole tile time to publish and the wait is normally already satisfied.
        static_assert(EPI != EPI_MAXOUT3_LN || HN == 96, "fused LayerNorm epilogue: 192-column tiles (64 units)");
        LnTile cur;
        cur.m0 = m0; cur.ntile = tile % n_tiles;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          float v[48];
          tmem_ld_x16_nowait(t_base + cc * 48, v);
          tmem_ld_x16_nowait(t_base + cc * 48 + 16, v + 16);
          tmem_ld_x16_nowait(t_base + cc * 48 + 32, v + 32);
          tmem_ld_wait_regs<48>(v);
          cur.wb[cc] = 0u;
          if (p.bias) {                                           // 48 consecutive bias values, 16-byte smem loads
#pragma unroll
            for (int k = 0; k < 12; ++k) {
              const float4 bq = ld_shared_f4(bias_s + n0 + cc * 48 + 4 * k);
              v[4 * k] += bq.x; v[4 * k + 1] += bq.y; v[4 * k + 2] += bq.z; v[4 * k + 3] += bq.w;
            }
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float a = v[3 * u], b = v[3 * u + 1], d = v[3 * u + 2];
            float best = a; uint32_t bi = 0;
            if (b > best) { best = b; bi = 1; }
            if (d > best) { best = d; bi = 2; }
            // round to bf16 first: the statistics are those of the values the two-kernel path stores
            const uint32_t hb = (uint32_t)__bfloat16_as_ushort(f2bf(best));
            const float hf = __uint_as_float(hb << 16);
            s1 += hf; s2 += hf * hf;
            if (u & 1) cur.hp[cc * 8 + u / 2] |= hb << 16; else cur.hp[cc * 8 + u / 2] = hb;
            cur.wb[cc] |= bi << (2 * u);
          }
        }
        // the accumulator stage is free again: the MMAs of the tile after next may start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && cta_rank != 0) mbar_arrive_cluster(map_to_cta(tmem_empty + acc, 0));
          else mbar_arrive(tmem_empty + acc);
        }
        // the two column halves of a row are in different warps of this CTA: add them up through shared
        // memory (double-buffered by tile parity), then ONE 16-byte store per row carrying this launch's tag
        {
          float2* xch = (float2*)(bias_s + p.N + 2 * (p.N / 3)) + (it & 1) * BM;
          if (half == 1) xch[q * 32 + lane] = make_float2(s1, s2);
          asm volatile("bar.sync 2, %0;" ::"n"(kNumThreads - 64) : "memory");
          if (half == 0) {
            const float2 o = xch[q * 32 + lane];
            const int srow = m0 + q * 32 + lane;                  // stats has m_groups * CL * 128 rows: always in range
            __stcg(p.ln.stats + (size_t)srow * n_tiles + cur.ntile,
                   make_float4(s1 + o.x, s2 + o.y, __uint_as_float(ln_tag), 0.f));
          }
        }
        if (ln_have) {
          // (issuing these loads before this tile's accumulator read-out was measured slower: the other
          // clusters have not published yet at that point and the slow re-read path is taken)
          ln_prefetch<BLOCK_N>(p, ln_prev, ln_ld, M, n_tiles, q, half, lane);
          ln_finish<BLOCK_N>(p, ln_prev, ln_ld, bias_s, M, n_tiles, q, half, lane, ln_tag);
        }
        ln_prev = cur; ln_have = true;
        ++it;
        continue;                                                 // tmem_empty was already signalled above
      } else {  // EPI_ATOMIC_F32: split-K partial sums reduced into the gradient buffer
        float* out = (float*)p.out;
#pragma unroll 1
        for (int c = 0; c < HN; c += 32) {
          float v[32];
          tmem_ld_x16_nowait(t_base + c, v);
          if (c + 16 < HN) tmem_ld_x16_nowait(t_base + c + 16, v + 16);
          tmem_ld_wait_regs<32>(v);
          if (row_ok) {
            float* dst = out + (size_t)row * p.ldo + n0 + c;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (c + i < HN && n0 + c + i < p.N) red_add_v4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && cta_rank != 0) mbar_arrive_cluster(map_to_cta(tmem_empty + acc, 0));   // the leader issues the MMAs
        else mbar_arrive(tmem_empty + acc);
      }
      ++it;
    }
    if (EPI == EPI_MAXOUT3_LN && ln_have) {
      ln_prefetch<BLOCK_N>(p, ln_prev, ln_ld, M, n_tiles, q, half, lane);
      ln_finish<BLOCK_N>(p, ln_prev, ln_ld, bias_s, M, n_tiles, q, half, lane, ln_tag);
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();      // no CTA leaves while a peer may still write to it
  if (warp == 1) { if (PAIR) tmem_dealloc_pair<C::kTmemCols>(tmem_base); else tmem_dealloc<C::kTmemCols>(tmem_base); }
  if (EPI == EPI_MAXOUT3_LN && threadIdx.x == 0) {
    // every CTA read the tag before it got here: the last one to finish arms the next launch
    __threadfence();
    if (atomicAdd(p.ln.seq + 1, 1u) == gridDim.x - 1u) {
      p.ln.seq[1] = 0u;
      p.ln.seq[0] = ln_tag + 1u == 0u ? 1u : ln_tag + 1u;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int BLOCK_N, int MODE, int EPI, int CL, bool PAIR = false, bool HALO = false, bool BRES = false>
static cudaError_t launch_one(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int num_sms, cudaStream_t s) {
  using C = Cfg<BLOCK_N, PAIR, HALO, BRES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BLOCK_N, MODE, EPI, CL, PAIR, HALO, BRES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int m_groups = (m_tiles + CL - 1) / CL;
  const int items = m_groups * n_tiles * (p.splits > 0 ? p.splits : 1);
  int clusters = num_sms / CL;
  if (items < clusters) clusters = items;
  if (BRES) clusters = clusters / n_tiles * n_tiles;      // every cluster keeps ONE N tile (see Cfg)
  if (clusters <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // launch.h
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_kernel<BLOCK_N, MODE, EPI, CL, PAIR, HALO, BRES>, a, b, p);
}

int gemm_block_k() { return BK; }

bool gemm_supports_cluster(int block_n, int mode, int epi) {
  if (mode == MODE_KK && (epi == EPI_STORE) && (block_n == 256 || block_n == 192 || block_n == 128)) return true;
  if (mode == MODE_KK && (epi == EPI_MAXOUT3 || epi == EPI_MAXOUT3_LN) && block_n == 192) return true;
  if (mode == MODE_MNMN && epi == EPI_ATOMIC_F32 && (block_n == 256 || block_n == 128)) return true;
  if (mode == MODE_KMN && epi == EPI_STORE && (block_n == 256 || block_n == 128)) return true;
  return false;
}

bool gemm_supports_halo(int block_n, int mode, int epi, int cluster) {
  if (cluster != 1 && cluster != 3) return false;
  if (mode == MODE_KK && epi == EPI_MAXOUT3_LN && block_n == 192) return cluster == 3;
  if (mode == MODE_KK && epi == EPI_MAXOUT3 && block_n == 192) return true;
  if (mode == MODE_KMN && epi == EPI_STORE && (block_n == 256 || block_n == 128)) return true;
  return false;
}

cudaError_t launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int block_n, int mode, int epi,
                        int cluster, int num_sms, cudaStream_t s) {
  if (epi == EPI_MAXOUT3_LN) {       // compiled for the pair-MMA cluster only (b200_ops falls back otherwise)
    if (block_n != 192 || mode != MODE_KK || cluster != 3) return cudaErrorInvalidValue;
    if (p.halo) return launch_one<192, MODE_KK, EPI_MAXOUT3_LN, 2, true, true>(a, b, p, num_sms, s);
    return launch_one<192, MODE_KK, EPI_MAXOUT3_LN, 2, true, false>(a, b, p, num_sms, s);
  }
  // window forward with the weights resident in shared memory (Cfg): K <= 256, bias (+ LayerNorm vectors) in 8 KB.
  // Measured at parity / 0.3 % slower than re-fetching B per tile (1.3172 vs 1.3132 ms per flagship step,
  // 72 instead of 74 clusters): the weight traffic is NOT what holds the tensor pipe at 68 % - the 4-k-block
  // mainloop (2.5 us) is as short as the epilogue of the previous tile.  Off unless SRB_GEMM_BRES=1.
  static const bool bres_on = std::getenv("SRB_GEMM_BRES") && std::getenv("SRB_GEMM_BRES")[0] == '1';
  if (p.halo && bres_on && cluster == 3 && block_n == 192 && mode == MODE_KK && p.K <= 256 && p.N >= 192 &&
      (num_sms / 2) >= p.N / 192) {
    if (epi == EPI_MAXOUT3 && p.N <= 2048)
      return launch_one<192, MODE_KK, EPI_MAXOUT3, 2, true, true, true>(a, b, p, num_sms, s);
  }
  if (p.halo) {
#define SRB_HALO(BN, MD, EP)                                                                       \
  if (block_n == BN && mode == MD && epi == EP) {                                                  \
    if (cluster == 3) return launch_one<BN, MD, EP, 2, true, true>(a, b, p, num_sms, s);           \
    return launch_one<BN, MD, EP, 1, false, true>(a, b, p, num_sms, s);                            \
  }
    SRB_HALO(192, MODE_KK, EPI_MAXOUT3)
    SRB_HALO(256, MODE_KMN, EPI_STORE)
    SRB_HALO(128, MODE_KMN, EPI_STORE)
#undef SRB_HALO
    return cudaErrorInvalidValue;
  }
#define SRB_CASE(BN, MD, EP)                                                              \
  if (block_n == BN && mode == MD && epi == EP) {                                         \
    if (cluster == 3) return launch_one<BN, MD, EP, 2, true>(a, b, p, num_sms, s);        \
    if (cluster == 2) return launch_one<BN, MD, EP, 2>(a, b, p, num_sms, s);              \
    return launch_one<BN, MD, EP, 1>(a, b, p, num_sms, s);                                \
  }
#define SRB_CASE1(BN, MD, EP) \
  if (block_n == BN && mode == MD && epi == EP) return launch_one<BN, MD, EP, 1>(a, b, p, num_sms, s);
  SRB_CASE(256, MODE_KK, EPI_STORE)
  SRB_CASE(192, MODE_KK, EPI_STORE)
  SRB_CASE(128, MODE_KK, EPI_STORE)
  SRB_CASE1(64, MODE_KK, EPI_STORE)
  SRB_CASE(192, MODE_KK, EPI_MAXOUT3)
  SRB_CASE1(96, MODE_KK, EPI_MAXOUT3)
  SRB_CASE1(96, MODE_KK, EPI_STORE)
  SRB_CASE(256, MODE_MNMN, EPI_ATOMIC_F32)
  SRB_CASE(128, MODE_MNMN, EPI_ATOMIC_F32)
  SRB_CASE1(64, MODE_MNMN, EPI_ATOMIC_F32)
  SRB_CASE1(256, MODE_KK, EPI_ATOMIC_F32)
  SRB_CASE1(128, MODE_KK, EPI_ATOMIC_F32)
  SRB_CASE1(64, MODE_KK, EPI_ATOMIC_F32)
  SRB_CASE(256, MODE_KMN, EPI_STORE)
  SRB_CASE1(192, MODE_KMN, EPI_STORE)
  SRB_CASE(128, MODE_KMN, EPI_STORE)
  SRB_CASE1(64, MODE_KMN, EPI_STORE)
#undef SRB_CASE
#undef SRB_CASE1
  return cudaErrorInvalidValue;
}

}  // namespace srb
