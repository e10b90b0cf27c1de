// K7: the BILUO (NER) transition loop as ONE kernel for the whole batch.
//
// Upstream (spaCy parser_model.pyx + _parser_internals) drives this loop from the
// CPU: one BLAS call + one C++ state update per transition step.  Here each doc
// gets a warp that walks its own state machine start to finish:
//   per step: sum the nF=3 precomputed feature rows (Yf) -> +bias -> maxout(nP=2)
//             -> upper layer (W_u staged transposed in smem) -> validity mask from the
//             state -> arg-max + masked softmax -> oracle gold action -> d_scores, loss
//             -> advance by the predicted action.
// The only serial dependency between steps is the tiny state, so nothing on the
// critical path waits on global memory: gold actions / step scales are preloaded into
// lane registers, and every feature row the NEXT step can possibly need (row i+1 slot 0,
// row i slots 1 and 2) is prefetched while the current step computes.
// The step records (feature rows, winning pieces, hidden, d_scores) are written once,
// so the whole backward pass is three batched GEMMs + one scatter kernel.
// Semantics are specified by models/transitions.py (BiluoSystem.batch_*) and
// models/transition_model.py::_biluo_steps_reference, which the tests diff against.
#include <cstdlib>

#include "common.cuh"
#include "launch.h"
#include "kernels.h"
#include "transition_common.cuh"

namespace srb {

constexpr int kWarpsPerBlock = 4;

// action a: 0 = OUT; a>0: kind = (a-1)%4 in {B,I,L,U}, label = (a-1)/4
__device__ __forceinline__ bool biluo_valid(int a, int ent_label, bool is_open, bool not_last) {
  if (a == 0) return !is_open;
  const int kind = (a - 1) & 3, lab = (a - 1) >> 2;
  if (!is_open) return kind == 3 || (kind == 0 && not_last);
  return lab == ent_label && (kind == 2 || (kind == 1 && not_last));
}

template <int PPL>
__device__ __forceinline__ void load_slot(const __nv_bfloat16* __restrict__ Yf, int row, int slot, int nOP, int lane,
                                          float out[PPL]) {
  load_bf16_vec<PPL>(Yf + ((size_t)row * 3 + slot) * nOP + lane * PPL, out);
}

// NJ = ceil(nA / 32) actions per lane, PPL = nO*nP/32 pre-activations per lane (nP == 2).
template <int NJ, int PPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) biluo_steps_kernel(BiluoArgs A) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int UPL = PPL / 2;
  const int nO = A.nO, nOP = A.nO * 2, nA = A.nA;
  float4* Wu4 = (float4*)smem_raw;                                // [nO/4][nA_pad] x float4
  float* bu_s = (float*)smem_raw + (size_t)nO * A.nA_pad;         // [nA_pad]
  float* hid_s = bu_s + A.nA_pad;                                 // [warps][nO]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_upper_weights(Wu4, bu_s, (const __nv_bfloat16*)A.Wu, (const __nv_bfloat16*)A.bu, nO, nA, A.nA_pad);
  __syncthreads();

  const int d = blockIdx.x * kWarpsPerBlock + warp;
  if (d >= A.B) return;
  const int n = A.doc_lens[d];
  const int row0 = A.doc_starts[d];
  const int tok0 = A.tok_off[d];
  const __nv_bfloat16* Yf = (const __nv_bfloat16*)A.Yf;
  float* hid_w = hid_s + warp * nO;
  // per-lane constants: bias and the two pad vectors for this lane's pre-activations
  float bias_r[PPL], pad12[PPL];
#pragma unroll
  for (int k = 0; k < PPL; ++k) {
    bias_r[k] = bf2f(((const __nv_bfloat16*)A.b)[lane * PPL + k]);
    pad12[k] = bf2f(((const __nv_bfloat16*)A.pad)[1 * nOP + lane * PPL + k]) +
               bf2f(((const __nv_bfloat16*)A.pad)[2 * nOP + lane * PPL + k]);
  }
  // gold actions and per-step scales of the first 64 steps live in lane registers
  const bool have_gold = A.gold != nullptr;
  int g_lo = -2, g_hi = -2;
  float sc_lo = 0.f, sc_hi = 0.f;
  if (have_gold) {
    g_lo = lane < n ? A.gold[tok0 + lane] : -1;
    g_hi = lane + 32 < n ? A.gold[tok0 + 32 + lane] : -1;
  }
  if (A.train) {
    sc_lo = lane < n ? A.inv_active[lane] : 0.f;
    sc_hi = lane + 32 < n ? A.inv_active[lane + 32] : 0.f;
  }
  int ent_start = -1, ent_label = -1;
  bool ent_ok = false;
  float loss_acc = 0.f;
  // feature vectors: nx0 = slot 0 of the current row; e1 = slot 1 of the open entity's first
  // row; l2 = slot 2 of the previous row (only meaningful while an entity is open)
  float nx0[PPL], e1[PPL], l2[PPL];
#pragma unroll
  for (int k = 0; k < PPL; ++k) { e1[k] = 0.f; l2[k] = 0.f; }
  if (n > 0) load_slot<PPL>(Yf, row0, 0, nOP, lane, nx0);

  for (int i = 0; i < n; ++i) {
    const bool is_open = ent_start >= 0;
    const bool not_last = (i + 1) < n;
    const int f0 = row0 + i;
    // prefetch everything the next step can need; none of it depends on this step's action
    float pn0[PPL], c1[PPL], c2[PPL];
    if (not_last) {
      load_slot<PPL>(Yf, f0 + 1, 0, nOP, lane, pn0);
      load_slot<PPL>(Yf, f0, 1, nOP, lane, c1);
      load_slot<PPL>(Yf, f0, 2, nOP, lane, c2);
    }
    // ---- hidden = maxout(b + slot0 + (open ? e1 + l2 : pad1 + pad2)) -----------------------
    float pre[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k) pre[k] = bias_r[k] + nx0[k] + (is_open ? (e1[k] + l2[k]) : pad12[k]);
    const size_t tok = (size_t)tok0 + i;
    float best_u[UPL];
    uint8_t which_u[UPL];
#pragma unroll
    for (int u = 0; u < UPL; ++u) {
      float best = pre[2 * u];
      int bi = 0;
      if (pre[2 * u + 1] > best) { best = pre[2 * u + 1]; bi = 1; }
      best_u[u] = best; which_u[u] = (uint8_t)bi;
      hid_w[lane * UPL + u] = best;
    }
    if (A.train) store_hidden_record<UPL>(A.which + tok * nO + lane * UPL, (__nv_bfloat16*)A.hid + tok * nO + lane * UPL,
                                          best_u, which_u);
    __syncwarp();
    // ---- upper layer: one LDS.128 of weights per four FMAs (transition_common.cuh) --------
    float sc[NJ];
    upper_layer<NJ>(Wu4, bu_s, hid_w, nO, A.nA_pad, lane, sc);
    bool ok[NJ];
    float mx = -3.0e38f;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int a = lane + 32 * j;
      ok[j] = a < nA && biluo_valid(a, ent_label, is_open, not_last);
      if (ok[j] && sc[j] > mx) { mx = sc[j]; arg = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    int g = -2;
    if (have_gold) {
      if (i < 64) g = __shfl_sync(0xffffffffu, i < 32 ? g_lo : g_hi, i & 31);
      else g = A.gold[tok];
    }
    int ga = -1;
    if (A.train) {
      // oracle: a single zero-cost action, or -1 = every valid action is zero-cost
      if (have_gold && g >= 0) {
        const int gk = g > 0 ? ((g - 1) & 3) : -1, gl = g > 0 ? ((g - 1) >> 2) : -1;
        if (!is_open) ga = (gk == -1 || gk == 0 || gk == 3) ? g : 0;
        else if (ent_ok && (gk == 1 || gk == 2) && gl == ent_label) ga = g;
      }
      if (ga >= 0 && !biluo_valid(ga, ent_label, is_open, not_last)) ga = -1;
      float e[NJ];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) { e[j] = ok[j] ? __expf(sc[j] - mx) : 0.f; sum += e[j]; }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      float scale;
      if (i < 64) scale = __shfl_sync(0xffffffffu, i < 32 ? sc_lo : sc_hi, i & 31);
      else scale = A.inv_active[i];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int a = lane + 32 * j;
        if (a < A.nA_pad) {
          float dv = 0.f;
          if (ga >= 0 && ok[j]) dv = (e[j] * inv - (a == ga ? 1.f : 0.f)) * scale;
          loss_acc += dv * dv;
          ((__nv_bfloat16*)A.d_scores)[tok * A.ld_scores + a] = f2bf(dv);
        }
      }
      if (lane < 3) {
        const int f1 = is_open ? row0 + ent_start : -1;
        A.feats[tok * 3 + lane] = lane == 0 ? f0 : (lane == 1 ? f1 : (is_open ? f0 - 1 : -1));
      }
    }
    if (lane == 0) A.actions[tok] = arg;
    // ---- advance by the predicted action (teacher forcing: by the oracle's, when it names one) --
    if (A.teacher && ga >= 0) arg = ga;
    const int kind = arg > 0 ? ((arg - 1) & 3) : -1;
    if (kind == 0) {
      ent_start = i; ent_label = (arg - 1) >> 2; ent_ok = (g == arg);
#pragma unroll
      for (int k = 0; k < PPL; ++k) e1[k] = c1[k];
    } else if (kind == 1) {
      ent_ok = ent_ok && (g == arg);
    } else {
      ent_start = -1; ent_label = -1; ent_ok = false;
    }
#pragma unroll
    for (int k = 0; k < PPL; ++k) { nx0[k] = pn0[k]; l2[k] = c2[k]; }
    __syncwarp();
  }
  if (A.train) {
    loss_acc = warp_sum(loss_acc);
    if (lane == 0 && loss_acc != 0.f) atomicAdd(A.loss, loss_acc);
  }
}

// ------------------------------------------------------------------------------------------
// Block-per-doc variant (round 2).  The warp-per-doc kernel above is bound by the dependent
// instruction chain of one step (~2.7 us: each lane evaluates 3 actions x 64 FMAs out of shared
// memory, then 15 shuffle rounds) times the longest doc, on a GPU that is 93 % idle meanwhile.
// Here a doc gets a CTA of NT threads and the step is cut across them:
//   thread p < nO*nP : one pre-activation  (slot rows are one coalesced bf16 each per thread)
//   thread a < nA    : one action score    (64 FMAs against ITS row of W_u, bf16 in shared memory,
//                                           16-byte loads, 4 accumulation chains)
// arg-max / soft-max run as warp shuffles + a 4-entry shared-memory combine.  Three CTA barriers
// per step (~50 cycles each) buy a ~5x shorter chain.  <= 64 registers: 8 CTAs per SM, so all
// 1024 docs of a flagship batch are resident at once.
// ------------------------------------------------------------------------------------------
constexpr int kBlkThreads = 128;
constexpr int kWuStride = 72;        // bf16 elements per W_u row in smem: 144 B = conflict-free 16-byte row reads

template <int NT>
__global__ void __launch_bounds__(NT, 8) biluo_block_kernel(BiluoArgs A) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nO = A.nO, nOP = A.nO * 2, nA = A.nA;
  __nv_bfloat16* Wu_s = (__nv_bfloat16*)smem_raw;                       // [nA_pad][kWuStride]
  float* hid_s = (float*)(Wu_s + (size_t)A.nA_pad * kWuStride);          // [2][nO] double-buffered
  float* red_v = hid_s + 2 * nO;                                         // [NT/32] warp maxima
  float* red_s = red_v + NT / 32;                                        // [NT/32] warp partial sums
  int* red_i = (int*)(red_s + NT / 32);                                  // [NT/32] warp arg-maxima
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = blockIdx.x;
  const int n = A.doc_lens[d];
  if (n <= 0) return;
  const int row0 = A.doc_starts[d];
  const int tok0 = A.tok_off[d];
  const __nv_bfloat16* Yf = (const __nv_bfloat16*)A.Yf;
  const __nv_bfloat16* Wu = (const __nv_bfloat16*)A.Wu;
  // stage W_u (bf16, padded rows)
  for (int i = tid; i < A.nA_pad * (nO / 8); i += NT) {
    const int a = i / (nO / 8), v = i % (nO / 8);
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (a < nA) w = *(const uint4*)(Wu + (size_t)a * nO + v * 8);
    *(uint4*)(Wu_s + (size_t)a * kWuStride + v * 8) = w;
  }
  const bool is_pre = tid < nOP, is_act = tid < A.nA_pad;
  const float bias_r = is_pre ? bf2f(((const __nv_bfloat16*)A.b)[tid]) : 0.f;
  const float pad12 = is_pre ? bf2f(((const __nv_bfloat16*)A.pad)[1 * nOP + tid]) + bf2f(((const __nv_bfloat16*)A.pad)[2 * nOP + tid]) : 0.f;
  const float bu_r = (tid < nA) ? bf2f(((const __nv_bfloat16*)A.bu)[tid]) : 0.f;
  const bool have_gold = A.gold != nullptr;
  int ent_start = -1, ent_label = -1;
  bool ent_ok = false;
  float loss_acc = 0.f;
  float nx0 = 0.f, e1 = 0.f, l2 = 0.f;
  if (is_pre) nx0 = bf2f(Yf[((size_t)row0 * 3 + 0) * nOP + tid]);
  int g_next = have_gold ? A.gold[tok0] : -2;
  float scale_next = A.train ? A.inv_active[0] : 0.f;
  __syncthreads();

  for (int i = 0; i < n; ++i) {
    const bool is_open = ent_start >= 0;
    const bool not_last = (i + 1) < n;
    const int f0 = row0 + i;
    const size_t tok = (size_t)tok0 + i;
    const int g = g_next;
    const float scale = scale_next;
    // prefetch everything the next step can need; none of it depends on this step's action
    float pn0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (not_last) {
      if (is_pre) {
        pn0 = bf2f(Yf[((size_t)(f0 + 1) * 3 + 0) * nOP + tid]);
        c1 = bf2f(Yf[((size_t)f0 * 3 + 1) * nOP + tid]);
        c2 = bf2f(Yf[((size_t)f0 * 3 + 2) * nOP + tid]);
      }
      if (have_gold) g_next = A.gold[tok + 1];
      if (A.train) scale_next = A.inv_active[i + 1];
    }
    // ---- hidden = maxout(b + slot0 + (open ? e1 + l2 : pad1 + pad2)): pieces sit in adjacent threads
    const float pre = bias_r + nx0 + (is_open ? (e1 + l2) : pad12);
    const float other = __shfl_xor_sync(0xffffffffu, pre, 1);
    float* hid_w = hid_s + (i & 1) * nO;
    if (is_pre && !(tid & 1)) {
      const int bi = other > pre ? 1 : 0;
      const float best = bi ? other : pre;
      hid_w[tid >> 1] = best;
      if (A.train) {
        A.which[tok * nO + (tid >> 1)] = (uint8_t)bi;
        ((__nv_bfloat16*)A.hid)[tok * nO + (tid >> 1)] = f2bf(best);
      }
    }
    __syncthreads();
    // ---- upper layer: my action's row of W_u (bf16, 16-byte loads) x the hidden vector (broadcast)
    float sc = -3.0e38f;
    bool ok = false;
    if (is_act) {
      float acc[4] = {bu_r, 0.f, 0.f, 0.f};
      const uint4* wrow = (const uint4*)(Wu_s + (size_t)tid * kWuStride);
#pragma unroll 4
      for (int v = 0; v < nO / 8; ++v) {
        const uint4 w = wrow[v];
        const float4 h0 = *(const float4*)(hid_w + v * 8), h1 = *(const float4*)(hid_w + v * 8 + 4);
        acc[0] = fmaf(h0.x, __uint_as_float(w.x << 16), acc[0]);
        acc[1] = fmaf(h0.y, __uint_as_float(w.x & 0xFFFF0000u), acc[1]);
        acc[2] = fmaf(h0.z, __uint_as_float(w.y << 16), acc[2]);
        acc[3] = fmaf(h0.w, __uint_as_float(w.y & 0xFFFF0000u), acc[3]);
        acc[0] = fmaf(h1.x, __uint_as_float(w.z << 16), acc[0]);
        acc[1] = fmaf(h1.y, __uint_as_float(w.z & 0xFFFF0000u), acc[1]);
        acc[2] = fmaf(h1.z, __uint_as_float(w.w << 16), acc[2]);
        acc[3] = fmaf(h1.w, __uint_as_float(w.w & 0xFFFF0000u), acc[3]);
      }
      sc = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      ok = tid < nA && biluo_valid(tid, ent_label, is_open, not_last);
    }
    // ---- arg-max over the valid actions: warp shuffles, then a 4-entry combine -------------------
    float mx = ok ? sc : -3.0e38f;
    int arg = ok ? tid : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    if (lane == 0) { red_v[warp] = mx; red_i[warp] = arg; }
    __syncthreads();
    mx = red_v[0]; arg = red_i[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) {
      const float om = red_v[w];
      const int oa = red_i[w];
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    int ga = -1;
    if (A.train) {
      // oracle: a single zero-cost action, or -1 = every valid action is zero-cost
      if (have_gold && g >= 0) {
        const int gk = g > 0 ? ((g - 1) & 3) : -1, gl = g > 0 ? ((g - 1) >> 2) : -1;
        if (!is_open) ga = (gk == -1 || gk == 0 || gk == 3) ? g : 0;
        else if (ent_ok && (gk == 1 || gk == 2) && gl == ent_label) ga = g;
      }
      if (ga >= 0 && !biluo_valid(ga, ent_label, is_open, not_last)) ga = -1;
      const float e = ok ? __expf(sc - mx) : 0.f;
      float sum = warp_sum(e);
      if (lane == 0) red_s[warp] = sum;       // (its readers of the previous step are behind two barriers)
      __syncthreads();
      sum = red_s[0];
#pragma unroll
      for (int w = 1; w < NT / 32; ++w) sum += red_s[w];
      if (is_act) {
        float dv = 0.f;
        if (ga >= 0 && ok) dv = (e / sum - (tid == ga ? 1.f : 0.f)) * scale;
        loss_acc += dv * dv;
        ((__nv_bfloat16*)A.d_scores)[tok * A.ld_scores + tid] = f2bf(dv);
      }
      if (tid < 3) {
        const int f1 = is_open ? row0 + ent_start : -1;
        A.feats[tok * 3 + tid] = tid == 0 ? f0 : (tid == 1 ? f1 : (is_open ? f0 - 1 : -1));
      }
    }
    if (tid == 0) A.actions[tok] = arg;
    // ---- advance by the predicted action (teacher forcing: by the oracle's, when it names one) --
    if (A.teacher && ga >= 0) arg = ga;
    const int kind = arg > 0 ? ((arg - 1) & 3) : -1;
    if (kind == 0) {
      ent_start = i; ent_label = (arg - 1) >> 2; ent_ok = (g == arg);
      e1 = c1;
    } else if (kind == 1) {
      ent_ok = ent_ok && (g == arg);
    } else {
      ent_start = -1; ent_label = -1; ent_ok = false;
    }
    nx0 = pn0; l2 = c2;
    // (red_v / red_i are rewritten only after the next step's first barrier; hid_s is double-buffered)
  }
  if (A.train) {
    loss_acc = warp_sum(loss_acc);
    if (lane == 0 && loss_acc != 0.f) atomicAdd(A.loss, loss_acc);
  }
}

static bool launch_biluo_block(const BiluoArgs& a, cudaStream_t s) {
  // shapes the block kernel covers: one thread per pre-activation and per action
  if (a.nP != 2 || a.nO * 2 > kBlkThreads || a.nA_pad > kBlkThreads || a.nO % 8 != 0 || a.nO > kWuStride) return false;
  const size_t smem = sizeof(__nv_bfloat16) * (size_t)a.nA_pad * kWuStride + sizeof(float) * (2 * a.nO + 2 * (kBlkThreads / 32)) +
                      sizeof(int) * (kBlkThreads / 32);
  launch_k(biluo_block_kernel<kBlkThreads>, a.B, kBlkThreads, smem, s, a);
  return true;
}

template <int NJ>
static void launch_nj(const BiluoArgs& a, int blocks, size_t smem, cudaStream_t s) {
  const int ppl = a.nO * 2 / 32;
#define SRB_PPL(P)                                                                                              \
  if (ppl == P) {                                                                                               \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(biluo_steps_kernel<NJ, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    launch_k(biluo_steps_kernel<NJ, P>, blocks, kWarpsPerBlock * 32, smem, s, a);                                     \
    return;                                                                                                     \
  }
  SRB_PPL(2) SRB_PPL(4) SRB_PPL(8)
#undef SRB_PPL
}

void launch_biluo_steps(BiluoArgs a, cudaStream_t s) {
  if (a.B <= 0) return;
  static const bool use_block = !(getenv("SRB_BILUO_BLOCK") && getenv("SRB_BILUO_BLOCK")[0] == '0');
  if (use_block && launch_biluo_block(a, s)) return;
  size_t smem = sizeof(float) * ((size_t)a.nO * a.nA_pad + a.nA_pad + kWarpsPerBlock * a.nO);
  int blocks = (a.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int nj = (a.nA_pad + 31) / 32;
  switch (nj) {
    case 1: launch_nj<1>(a, blocks, smem, s); break;
    case 2: launch_nj<2>(a, blocks, smem, s); break;
    case 3: launch_nj<3>(a, blocks, smem, s); break;
    case 4: launch_nj<4>(a, blocks, smem, s); break;
    case 5: launch_nj<5>(a, blocks, smem, s); break;
    case 6: launch_nj<6>(a, blocks, smem, s); break;
    case 7: launch_nj<7>(a, blocks, smem, s); break;
    default: launch_nj<8>(a, blocks, smem, s); break;
  }
}

// Backward scatter: one warp per recorded step.
__global__ void __launch_bounds__(128) transition_scatter_kernel(const __nv_bfloat16* __restrict__ d_hid,
                                                                 const uint8_t* __restrict__ which,
                                                                 const int32_t* __restrict__ feats,
                                                                 float* __restrict__ dYf, float* __restrict__ dpad,
                                                                 float* __restrict__ db, int S, int nF, int nO,
                                                                 int nP) {
  pdl_prologue();
  extern __shared__ float sacc[];          // [nF+1][nOP]: dpad rows then db
  const int nOP = nO * nP;
  for (int i = threadIdx.x; i < (nF + 1) * nOP; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = blockIdx.x * 4 + warp; s < S; s += gridDim.x * 4) {
    for (int o = lane; o < nO; o += 32) {
      const float v = bf2f(d_hid[(size_t)s * nO + o]);
      if (v == 0.f) continue;
      const int j = o * nP + which[(size_t)s * nO + o];
      atomicAdd(&sacc[nF * nOP + j], v);
      for (int f = 0; f < nF; ++f) {
        const int row = feats[(size_t)s * nF + f];
        if (row >= 0) atomicAdd(dYf + ((size_t)row * nF + f) * nOP + j, v);
        else atomicAdd(&sacc[f * nOP + j], v);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nF * nOP; i += blockDim.x) if (sacc[i] != 0.f) atomicAdd(dpad + i, sacc[i]);
  for (int i = threadIdx.x; i < nOP; i += blockDim.x) if (sacc[nF * nOP + i] != 0.f) atomicAdd(db + i, sacc[nF * nOP + i]);
}

void launch_transition_scatter(const void* d_hid, const uint8_t* which, const int32_t* feats, float* dYf,
                               float* dpad, float* db, int S, int nF, int nO, int nP, cudaStream_t s) {
  if (S <= 0) return;
  int blocks = (S + 3) / 4;
  if (blocks > 148 * 8) blocks = 148 * 8;
  size_t smem = sizeof(float) * (size_t)(nF + 1) * nO * nP;
  launch_k(transition_scatter_kernel, blocks, 128, smem, s, (const __nv_bfloat16*)d_hid, which, feats, dYf, dpad, db, S,
                                                     nF, nO, nP);
}

}  // namespace srb
