// K7 (parser): the arc-eager transition loop with a dynamic oracle, ONE kernel per batch.
//
// Upstream this is spaCy's Cython/C++ `_parser_internals` driven per step from the host.
// Here one warp owns one doc and runs its whole derivation on the device:
//   state (stack, heads, labels, leftmost/rightmost children, stack membership) lives in
//   shared memory; per step the warp
//     * builds the 8 feature tokens [B0, B1, S0, S1, S2, L(B0), L(S0), R(S0)],
//     * sums the 8 precomputed feature rows (Yf) + bias -> maxout -> upper layer (W_u^T in smem),
//     * derives the valid-action mask from the state,
//     * evaluates the dynamic oracle (Goldberg & Nivre 2012) warp-parallel: the two counts it
//       needs (gold children of B0 among headless stack items; gold children of S0 in the
//       buffer) are ballots over the doc's tokens,
//     * d_scores = softmax(valid) - softmax(valid & min-cost), loss, arg-max valid action,
//     * applies the predicted action.
// Records (feature rows, winning pieces, hidden, d_scores) are written per step so the backward
// pass is the same three batched GEMMs + scatter kernel as for NER.
// Spec: models/transitions.py::ArcEagerSystem and transition_model.py::_arc_steps_reference.
#include "common.cuh"
#include "launch.h"
#include "kernels.h"
#include "transition_common.cuh"

namespace srb {

constexpr int kArcWarps = 4;
constexpr int kArcSmemLimit = 200 * 1024;

// Per-warp parser state in shared memory, sized at launch for the longest doc of the batch
// (A.max_n, a multiple of 32): 7 int arrays + 1 byte array of max_n entries.
struct ArcWarpState {
  int* stack;
  int* heads;
  int* labels;
  int* lc;
  int* rc;
  int* gh;                         // gold head: -2 missing, -1 root, else doc-relative index
  int* gl;                         // gold label or -1
  unsigned char* in_stack;
};
__host__ __device__ inline size_t arc_state_bytes(int max_n) { return (size_t)max_n * (7 * sizeof(int) + 1); }
__device__ __forceinline__ ArcWarpState arc_state(unsigned char* base, int warp, int max_n) {
  unsigned char* p = base + (size_t)warp * arc_state_bytes(max_n);
  ArcWarpState S;
  S.stack = (int*)p; S.heads = S.stack + max_n; S.labels = S.heads + max_n; S.lc = S.labels + max_n;
  S.rc = S.lc + max_n; S.gh = S.rc + max_n; S.gl = S.gh + max_n;
  S.in_stack = (unsigned char*)(S.gl + max_n);
  return S;
}

template <int NP, int UPL, int NJ>
__global__ void __launch_bounds__(kArcWarps * 32) arc_eager_steps_kernel(ArcArgs A) {
  pdl_prologue();
  constexpr int PPL = NP * UPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nO = A.nO, nOP = A.nO * NP, nA = A.nA;
  float4* Wu4 = (float4*)smem_raw;                                // [nO/4][nA_pad] x float4
  float* bu_s = (float*)smem_raw + (size_t)nO * A.nA_pad;         // [nA_pad]
  float* pad_s = bu_s + A.nA_pad;                                 // [8][nOP]
  float* hid_s = pad_s + 8 * nOP;                                 // [warps][nO]
  unsigned char* state_base = (unsigned char*)(hid_s + kArcWarps * nO);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_upper_weights(Wu4, bu_s, (const __nv_bfloat16*)A.Wu, (const __nv_bfloat16*)A.bu, nO, nA, A.nA_pad);
  for (int i = threadIdx.x; i < 8 * nOP; i += blockDim.x) pad_s[i] = bf2f(((const __nv_bfloat16*)A.pad)[i]);
  __syncthreads();

  const int d = blockIdx.x * kArcWarps + warp;
  if (d >= A.B) return;
  const int n = min(A.doc_lens[d], A.max_n);       // the host sizes max_n for the longest doc of the batch
  const int row0 = A.doc_starts[d];
  const int tok0 = A.tok_off[d];
  const int rec0 = A.step_off[d];
  const ArcWarpState S = arc_state(state_base, warp, A.max_n);
  float* hid_w = hid_s + warp * nO;
  const bool have_gold = A.gold_heads != nullptr;
  for (int t = lane; t < n; t += 32) {
    S.heads[t] = -1; S.labels[t] = -1; S.lc[t] = -1; S.rc[t] = -1; S.in_stack[t] = 0;
    int g = -2, l = -1;
    if (have_gold) {
      const int h = A.gold_heads[tok0 + t];
      g = h < 0 ? -2 : (h == t ? -1 : h);
      l = A.gold_labels[tok0 + t];
    }
    S.gh[t] = g; S.gl[t] = l;
  }
  __syncwarp();
  float bias_r[PPL];
#pragma unroll
  for (int k = 0; k < PPL; ++k) bias_r[k] = bf2f(((const __nv_bfloat16*)A.b)[lane * PPL + k]);
  const __nv_bfloat16* Yf = (const __nv_bfloat16*)A.Yf;

  int sp = 0, b = 0, step = 0;
  float loss_acc = 0.f;
  const int max_steps = 2 * n;
  while (!(b >= n && sp <= 1) && step < max_steps) {
    const bool has_buf = b < n, has_stack = sp > 0;
    const int b0 = has_buf ? b : -1;
    const int b1 = (b + 1 < n) ? b + 1 : -1;
    const int s0 = sp > 0 ? S.stack[sp - 1] : -1;
    const int s1 = sp > 1 ? S.stack[sp - 2] : -1;
    const int s2 = sp > 2 ? S.stack[sp - 3] : -1;
    const int f[8] = {b0, b1, s0, s1, s2, b0 >= 0 ? S.lc[b0] : -1, s0 >= 0 ? S.lc[s0] : -1, s0 >= 0 ? S.rc[s0] : -1};
    const bool s0_headed = s0 >= 0 && S.heads[s0] >= 0;
    // ---- hidden ---------------------------------------------------------------------------
    // all eight feature rows are requested before the first one is used (missing features
    // read the doc's first row - valid memory - and are replaced by the pad vector below)
    float fv[8][PPL];
#pragma unroll
    for (int ff = 0; ff < 8; ++ff) {
      const int fi = f[ff] >= 0 ? f[ff] : 0;
      load_bf16_vec<PPL>(Yf + ((size_t)(row0 + fi) * 8 + ff) * nOP + lane * PPL, fv[ff]);
    }
    float pre[PPL];
#pragma unroll
    for (int k = 0; k < PPL; ++k) pre[k] = bias_r[k];
#pragma unroll
    for (int ff = 0; ff < 8; ++ff) {
      if (f[ff] >= 0) {
#pragma unroll
        for (int k = 0; k < PPL; ++k) pre[k] += fv[ff][k];
      } else {
#pragma unroll
        for (int k = 0; k < PPL; ++k) pre[k] += pad_s[ff * nOP + lane * PPL + k];
      }
    }
    const size_t rec = (size_t)rec0 + step;
    float best_u[UPL];
    uint8_t which_u[UPL];
#pragma unroll
    for (int u = 0; u < UPL; ++u) {
      float best = pre[u * NP];
      int bi = 0;
#pragma unroll
      for (int q = 1; q < NP; ++q) if (pre[u * NP + q] > best) { best = pre[u * NP + q]; bi = q; }
      best_u[u] = best; which_u[u] = (uint8_t)bi;
      hid_w[lane * UPL + u] = best;
    }
    if (A.train) store_hidden_record<UPL>(A.which + rec * nO + lane * UPL, (__nv_bfloat16*)A.hid + rec * nO + lane * UPL,
                                          best_u, which_u);
    __syncwarp();
    // ---- upper layer ----------------------------------------------------------------------
    float sc[NJ];
    upper_layer<NJ>(Wu4, bu_s, hid_w, nO, A.nA_pad, lane, sc);
    // ---- validity + arg-max ---------------------------------------------------------------
    bool ok[NJ];
    float mx = -3.0e38f;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int a = lane + 32 * j;
      bool v = false;
      if (a < nA) {
        if (a == 0) v = has_buf;
        else if (a == 1) v = has_stack && (s0_headed || !has_buf);
        else {
          const bool right = ((a - 2) & 1) != 0;
          v = has_stack && has_buf && (right || !s0_headed);
        }
      }
      ok[j] = v;
      if (v && sc[j] > mx) { mx = sc[j]; arg = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    int teach = -1;
    if (A.train) {
      // ---- dynamic oracle ------------------------------------------------------------------
      int cnt_stack_b0 = 0, cnt_buf_s0 = 0;
      if (have_gold) {
        for (int t = lane; t < n; t += 32) {
          const int g = S.gh[t];
          if (b0 >= 0 && S.in_stack[t] && S.heads[t] < 0 && g == b0) cnt_stack_b0 += 1;
          if (s0 >= 0 && t >= b && g == s0) cnt_buf_s0 += 1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          cnt_stack_b0 += __shfl_xor_sync(0xffffffffu, cnt_stack_b0, o);
          cnt_buf_s0 += __shfl_xor_sync(0xffffffffu, cnt_buf_s0, o);
        }
      }
      const int gb = b0 >= 0 ? S.gh[b0] : -2, gs = s0 >= 0 ? S.gh[s0] : -2;
      const bool gb_in_stack = gb >= 0 && S.in_stack[gb];
      int c_shift = (gb_in_stack ? 1 : 0) + cnt_stack_b0;
      int c_reduce = (s0 >= 0 && !s0_headed && !has_buf) ? 0 : cnt_buf_s0;
      int c_left = cnt_buf_s0 + ((gs != -2 && gs != b0 && (gs == -1 || gs > b0)) ? 1 : 0);
      int c_right = ((gb != -2 && gb != s0 && (gb == -1 || gb_in_stack || gb > b0)) ? 1 : 0) + cnt_stack_b0;
      const int gl_s0 = s0 >= 0 ? S.gl[s0] : -1, gl_b0 = b0 >= 0 ? S.gl[b0] : -1;
      int cost[NJ];
      int cmin = 1 << 20;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int a = lane + 32 * j;
        int c = 1 << 20;
        if (ok[j]) {
          if (!have_gold) c = 0;
          else if (a == 0) c = c_shift;
          else if (a == 1) c = c_reduce;
          else {
            const int lab = (a - 2) >> 1;
            if ((a - 2) & 1) c = c_right + ((gb == s0 && gl_b0 >= 0 && gl_b0 != lab) ? 1 : 0);
            else c = c_left + ((gs == b0 && gl_s0 >= 0 && gl_s0 != lab) ? 1 : 0);
          }
        }
        cost[j] = c;
        cmin = min(cmin, c);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
      if (A.teacher) {                 // first minimum-cost valid action (what the host loop follows too)
        int ga = 1 << 20;
#pragma unroll
        for (int j = 0; j < NJ; ++j) if (ok[j] && cost[j] == cmin) ga = min(ga, lane + 32 * j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ga = min(ga, __shfl_xor_sync(0xffffffffu, ga, o));
        if (ga < (1 << 20)) teach = ga;
      }
      float e[NJ], eg[NJ];
      float sum = 0.f, gsum = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        e[j] = ok[j] ? __expf(sc[j] - mx) : 0.f;
        eg[j] = (ok[j] && cost[j] == cmin) ? e[j] : 0.f;
        sum += e[j]; gsum += eg[j];
      }
      sum = warp_sum(sum); gsum = warp_sum(gsum);
      const float inv = 1.f / sum, ginv = gsum > 0.f ? 1.f / gsum : 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int a = lane + 32 * j;
        if (a < A.nA_pad) {
          float dv = 0.f;
          if (ok[j] && gsum > 0.f) dv = (e[j] * inv - eg[j] * ginv) * A.scale;
          loss_acc += dv * dv;
          ((__nv_bfloat16*)A.d_scores)[rec * A.ld_scores + a] = f2bf(dv);
        }
      }
      if (lane < 8) {
        const int fv = lane == 0 ? f[0] : lane == 1 ? f[1] : lane == 2 ? f[2] : lane == 3 ? f[3]
                     : lane == 4 ? f[4] : lane == 5 ? f[5] : lane == 6 ? f[6] : f[7];
        A.feats[rec * 8 + lane] = fv >= 0 ? row0 + fv : -1;
      }
    }
    if (teach >= 0) arg = teach;
    if (A.history && lane == 0) A.history[rec] = arg;
    // ---- apply the predicted (or teacher-forced) action ----------------------------------------
    __syncwarp();      // every lane has finished reading the state (oracle ballots) before lane 0 mutates it
    if (lane == 0) {
      if (arg == 0) { S.stack[sp] = b; S.in_stack[b] = 1; }
      else if (arg == 1) { S.in_stack[s0] = 0; }
      else {
        const int lab = (arg - 2) >> 1;
        if ((arg - 2) & 1) {            // RIGHT: S0 -> B0, push B0
          S.heads[b0] = s0; S.labels[b0] = lab;
          if (S.rc[s0] < b0) S.rc[s0] = b0;
          S.stack[sp] = b0; S.in_stack[b0] = 1;
        } else {                        // LEFT: B0 -> S0, pop S0
          S.heads[s0] = b0; S.labels[s0] = lab;
          if (S.lc[b0] < 0 || s0 < S.lc[b0]) S.lc[b0] = s0;
          S.in_stack[s0] = 0;
        }
      }
    }
    if (arg == 0) { sp += 1; b += 1; }
    else if (arg == 1) { sp -= 1; }
    else if ((arg - 2) & 1) { sp += 1; b += 1; }
    else { sp -= 1; }
    step += 1;
    __syncwarp();
  }
  for (int t = lane; t < n; t += 32) {
    A.heads_out[tok0 + t] = S.heads[t] >= 0 ? S.heads[t] : t;
    A.labels_out[tok0 + t] = S.labels[t];
  }
  if (lane == 0) A.n_steps[d] = step;
  if (A.train) {
    loss_acc = warp_sum(loss_acc);
    if (lane == 0 && loss_acc != 0.f) atomicAdd(A.loss, loss_acc);
  }
}

template <int NP, int UPL, int NJ>
static void launch_arc(const ArcArgs& a, int blocks, size_t smem, cudaStream_t s) {
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(arc_eager_steps_kernel<NP, UPL, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_k(arc_eager_steps_kernel<NP, UPL, NJ>, blocks, kArcWarps * 32, smem, s, a);
}

template <int NP, int UPL>
static void launch_arc_nj(const ArcArgs& a, int blocks, size_t smem, cudaStream_t s) {
  const int nj = (a.nA_pad + 31) / 32;
  switch (nj) {
    case 1: launch_arc<NP, UPL, 1>(a, blocks, smem, s); break;
    case 2: launch_arc<NP, UPL, 2>(a, blocks, smem, s); break;
    case 3: launch_arc<NP, UPL, 3>(a, blocks, smem, s); break;
    case 4: launch_arc<NP, UPL, 4>(a, blocks, smem, s); break;
    default: launch_arc<NP, UPL, 6>(a, blocks, smem, s); break;     // nA <= 192
  }
}

static size_t arc_fixed_smem(int nO, int nP, int nA_pad) {
  return sizeof(float) * ((size_t)nO * nA_pad + nA_pad + 8 * nO * nP + kArcWarps * nO);
}

bool launch_arc_eager_steps(ArcArgs a, cudaStream_t s) {
  if (a.B <= 0) return true;
  if (a.nO % 32 != 0 || a.nA_pad > 192) return false;
  const int upl = a.nO / 32;
  a.max_n = ((a.max_n > 0 ? a.max_n : 128) + 31) / 32 * 32;
  const size_t smem = arc_fixed_smem(a.nO, a.nP, a.nA_pad) + kArcWarps * arc_state_bytes(a.max_n);
  if (smem > (size_t)kArcSmemLimit) return false;            // docs this long: host state machine
  const int blocks = (a.B + kArcWarps - 1) / kArcWarps;
#define SRB_ARC(NP_, UPL_) \
  if (a.nP == NP_ && upl == UPL_) { launch_arc_nj<NP_, UPL_>(a, blocks, smem, s); return true; }
  SRB_ARC(2, 2) SRB_ARC(2, 4) SRB_ARC(3, 2) SRB_ARC(3, 4) SRB_ARC(2, 1) SRB_ARC(3, 1)
#undef SRB_ARC
  return false;
}

// Longest doc (tokens) the device kernel can hold for this head shape: the per-warp state has to fit
// in shared memory next to the staged upper-layer weights.
int arc_eager_max_doc_len(int nO, int nP, int nA) {
  const int nA_pad = (nA + 7) / 8 * 8;
  const size_t fixed = arc_fixed_smem(nO, nP, nA_pad);
  if (fixed >= (size_t)kArcSmemLimit) return 0;
  const size_t per_tok = kArcWarps * (7 * sizeof(int) + 1);
  return (int)(((size_t)kArcSmemLimit - fixed) / per_tok / 32 * 32);
}

}  // namespace srb
