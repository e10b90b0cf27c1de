// K7: the BILUO (NER) transition loop as ONE kernel for the whole batch.
//
// Upstream (spaCy parser_model.pyx + _parser_internals) drives this loop from the
// CPU: one BLAS call + one C++ state update per transition step.  Here each doc
// gets a warp that walks its own state machine start to finish:
//   per step: gather nF=3 precomputed feature rows (Yf) -> +bias -> maxout(nP=2)
//             -> upper layer (W_u staged transposed in smem) -> validity mask from the
//             state -> arg-max + masked softmax -> oracle gold action -> d_scores, loss
//             -> advance by the predicted action.
// The step records (feature rows, winning pieces, hidden, d_scores) are written
// once, so the whole backward pass is three batched GEMMs + one scatter kernel.
// Semantics are specified by models/transitions.py (BiluoSystem.batch_*) and
// models/transition_model.py::_biluo_steps_reference, which the tests diff against.
#include "common.cuh"
#include "kernels.h"

namespace srb {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxActionsPerLane = 8;      // nA <= 256

// action a: 0 = OUT; a>0: kind = (a-1)%4 in {B,I,L,U}, label = (a-1)/4
__device__ __forceinline__ bool biluo_valid(int a, int ent_label, bool is_open, bool not_last) {
  if (a == 0) return !is_open;
  const int kind = (a - 1) & 3, lab = (a - 1) >> 2;
  if (!is_open) return kind == 3 || (kind == 0 && not_last);
  return lab == ent_label && (kind == 2 || (kind == 1 && not_last));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) biluo_steps_kernel(BiluoArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nO = A.nO, nOP = A.nO * A.nP, nA = A.nA;
  float* WuT = (float*)smem_raw;                                  // [nO][nA_pad]
  float* bu_s = WuT + (size_t)nO * A.nA_pad;                      // [nA_pad]
  float* bias_s = bu_s + A.nA_pad;                                // [nOP]
  float* pad_s = bias_s + nOP;                                    // [3][nOP]
  float* hid_s = pad_s + 3 * nOP;                                 // [warps][nO]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* Wu = (const __nv_bfloat16*)A.Wu;
  for (int i = threadIdx.x; i < nO * A.nA_pad; i += blockDim.x) {
    const int o = i / A.nA_pad, a = i - o * A.nA_pad;
    WuT[i] = a < nA ? bf2f(Wu[(size_t)a * nO + o]) : 0.f;
  }
  for (int i = threadIdx.x; i < A.nA_pad; i += blockDim.x)
    bu_s[i] = i < nA ? bf2f(((const __nv_bfloat16*)A.bu)[i]) : 0.f;
  for (int i = threadIdx.x; i < nOP; i += blockDim.x) bias_s[i] = bf2f(((const __nv_bfloat16*)A.b)[i]);
  for (int i = threadIdx.x; i < 3 * nOP; i += blockDim.x) pad_s[i] = bf2f(((const __nv_bfloat16*)A.pad)[i]);
  __syncthreads();

  const int d = blockIdx.x * kWarpsPerBlock + warp;
  if (d >= A.B) return;
  const int n = A.doc_lens[d];
  const int row0 = A.doc_starts[d];
  const int tok0 = A.tok_off[d];
  const __nv_bfloat16* Yf = (const __nv_bfloat16*)A.Yf;
  float* hid_w = hid_s + warp * nO;
  // each lane owns pre-activations [lane*PPL, lane*PPL+PPL): with nP=2 that is whole units
  const int ppl = nOP / 32;            // 4 for nO=64,nP=2
  const int upl = nO / 32;             // 2
  int ent_start = -1, ent_label = -1;
  bool ent_ok = false;
  float loss_acc = 0.f;

  for (int i = 0; i < n; ++i) {
    const bool is_open = ent_start >= 0;
    const bool not_last = (i + 1) < n;
    const int f0 = row0 + i;
    const int f1 = is_open ? row0 + ent_start : -1;
    const int f2 = is_open ? f0 - 1 : -1;
    // ---- hidden = maxout(b + sum_f Yf[row_f, f] | pad[f]) -----------------------------
    float pre[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < ppl) pre[k] = bias_s[lane * ppl + k];
    {
      const __nv_bfloat16* p0 = Yf + ((size_t)f0 * 3 + 0) * nOP + lane * ppl;
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k < ppl) pre[k] += bf2f(p0[k]);
      if (f1 >= 0) {
        const __nv_bfloat16* p1 = Yf + ((size_t)f1 * 3 + 1) * nOP + lane * ppl;
        const __nv_bfloat16* p2 = Yf + ((size_t)f2 * 3 + 2) * nOP + lane * ppl;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < ppl) pre[k] += bf2f(p1[k]) + bf2f(p2[k]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < ppl) pre[k] += pad_s[1 * nOP + lane * ppl + k] + pad_s[2 * nOP + lane * ppl + k];
      }
    }
    const size_t tok = (size_t)tok0 + i;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < upl) {
        // nP == 2 (checked by the launcher): static register indices, no local-memory array
        float best = pre[2 * u];
        int bi = 0;
        if (pre[2 * u + 1] > best) { best = pre[2 * u + 1]; bi = 1; }
        const int o = lane * upl + u;
        hid_w[o] = best;
        if (A.train) {
          A.which[tok * nO + o] = (uint8_t)bi;
          ((__nv_bfloat16*)A.hid)[tok * nO + o] = f2bf(best);
        }
      }
    }
    __syncwarp();
    // ---- scores, validity, arg-max -----------------------------------------------------
    float sc[kMaxActionsPerLane];
    bool ok[kMaxActionsPerLane];
    float mx = -3.0e38f;
    int arg = 0;
    // upper layer: o-outer / action-inner so the per-lane accumulators are independent FMA chains
    const int nj = (nA + 31) >> 5;
#pragma unroll
    for (int j = 0; j < kMaxActionsPerLane; ++j) {
      const int a = lane + 32 * j;
      sc[j] = (j < nj && a < nA) ? bu_s[a] : 0.f;
    }
    for (int o = 0; o < nO; ++o) {
      const float h = hid_w[o];
      const float* wrow = WuT + o * A.nA_pad + lane;
#pragma unroll
      for (int j = 0; j < kMaxActionsPerLane; ++j)
        if (j < nj && lane + 32 * j < A.nA_pad) sc[j] = fmaf(h, wrow[32 * j], sc[j]);
    }
#pragma unroll
    for (int j = 0; j < kMaxActionsPerLane; ++j) {
      const int a = lane + 32 * j;
      ok[j] = false;
      if (a < nA) {
        ok[j] = biluo_valid(a, ent_label, is_open, not_last);
        if (ok[j] && (sc[j] > mx)) { mx = sc[j]; arg = a; }
      } else {
        sc[j] = -3.0e38f;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, mx, o);
      int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    const int g = A.gold ? A.gold[tok] : -2;
    if (A.train) {
      // oracle: a single zero-cost action, or -1 = every valid action is zero-cost
      int ga = -1;
      if (A.gold && g >= 0) {
        const int gk = g > 0 ? ((g - 1) & 3) : -1, gl = g > 0 ? ((g - 1) >> 2) : -1;
        if (!is_open) ga = (gk == -1 || gk == 0 || gk == 3) ? g : 0;
        else if (ent_ok && (gk == 1 || gk == 2) && gl == ent_label) ga = g;
      }
      if (ga >= 0 && !biluo_valid(ga, ent_label, is_open, not_last)) ga = -1;
      float sum = 0.f;
      float e[kMaxActionsPerLane];
#pragma unroll
      for (int j = 0; j < kMaxActionsPerLane; ++j) { e[j] = ok[j] ? __expf(sc[j] - mx) : 0.f; sum += e[j]; }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      const float scale = A.inv_active[i];
#pragma unroll
      for (int j = 0; j < kMaxActionsPerLane; ++j) {
        const int a = lane + 32 * j;
        if (a < A.nA_pad) {
          float dv = 0.f;
          if (ga >= 0 && ok[j]) dv = (e[j] * inv - (a == ga ? 1.f : 0.f)) * scale;
          loss_acc += dv * dv;
          ((__nv_bfloat16*)A.d_scores)[tok * A.nA_pad + a] = f2bf(dv);
        }
      }
      if (lane < 3) A.feats[tok * 3 + lane] = lane == 0 ? f0 : (lane == 1 ? f1 : f2);
    }
    if (lane == 0) A.actions[tok] = arg;
    // ---- advance by the predicted action ----------------------------------------------
    const int kind = arg > 0 ? ((arg - 1) & 3) : -1;
    if (kind == 0) { ent_start = i; ent_label = (arg - 1) >> 2; ent_ok = (g == arg); }
    else if (kind == 1) { ent_ok = ent_ok && (g == arg); }
    else { ent_start = -1; ent_label = -1; ent_ok = false; }
    __syncwarp();
  }
  if (A.train) {
    loss_acc = warp_sum(loss_acc);
    if (lane == 0 && loss_acc != 0.f) atomicAdd(A.loss, loss_acc);
  }
}

void launch_biluo_steps(BiluoArgs a, cudaStream_t s) {
  if (a.B <= 0) return;
  const int nOP = a.nO * a.nP;
  size_t smem = sizeof(float) * ((size_t)a.nO * a.nA_pad + a.nA_pad + nOP + 3 * nOP + kWarpsPerBlock * a.nO);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(biluo_steps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int blocks = (a.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  biluo_steps_kernel<<<blocks, kWarpsPerBlock * 32, smem, s>>>(a);
}

// Backward scatter: one warp per recorded step.
__global__ void __launch_bounds__(128) transition_scatter_kernel(const __nv_bfloat16* __restrict__ d_hid,
                                                                 const uint8_t* __restrict__ which,
                                                                 const int32_t* __restrict__ feats,
                                                                 float* __restrict__ dYf, float* __restrict__ dpad,
                                                                 float* __restrict__ db, int S, int nF, int nO,
                                                                 int nP) {
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
  transition_scatter_kernel<<<blocks, 128, smem, s>>>((const __nv_bfloat16*)d_hid, which, feats, dYf, dpad, db, S,
                                                     nF, nO, nP);
}

}  // namespace srb
