// Host/device shared declarations for the tcgen05 GEMM family.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "comm_launch.h"

namespace srb {

// MODE_KK:   A (M,K) and B (N,K) both K-major ("NT")           - forward / plain linear
// MODE_MNMN: A (K,M) and B (K,N) both MN-major ("TN")          - dW straight from (T,.) arrays
// MODE_KMN:  A (M,K) K-major, B (K,N) MN-major ("NN")          - dX = dY @ W with W as stored,
//            no per-step transpose of the weights; b_row_off = K (row) offset, b_col_off = N offset
enum { MODE_KK = 0, MODE_MNMN = 1, MODE_KMN = 2 };
enum { EPI_STORE = 0, EPI_MAXOUT3 = 1, EPI_ATOMIC_F32 = 2, EPI_MAXOUT3_LN = 3 };

// EPI_MAXOUT3_LN: maxout + LayerNorm + dropout + residual + mask fused into the GEMM epilogue.  A
// row's nO units are spread over n_tiles = 3*nO / 192 N tiles that DIFFERENT clusters compute at about
// the same time; what a LayerNorm needs from the other tiles is two numbers per row (sum, sum of
// squares).  Each epilogue thread keeps its 32 maxout outputs of the row in registers, publishes its
// partial sums (added to those of the row's other column half through shared memory) as ONE 16-byte
// store {sum, sumsq, launch tag, 0} into `stats`, and - one tile later, software-pipelined - reads the
// n_tiles entries of its row back from L2 (one round trip, re-read
// until every tag is this launch's), then normalises and stores its own units.  The activations never
// leave the SM between the accumulator and the final Y / xhat / which stores; there are no counters
// and no atomics on the data path.  The tag is a device-side launch sequence number (`seq[0]`), bumped
// by the last CTA to finish (`seq[1]` counts finished CTAs), so CUDA-graph replays need no host input.
struct LnArgs {
  const __nv_bfloat16* G;       // (nO) LayerNorm gain
  const __nv_bfloat16* beta;    // (nO)
  const __nv_bfloat16* xres;    // optional residual input (rows x nO, pitch ld_res)
  int ld_res;
  const float* mask;            // (rows) 0 = pad row: outputs are zero
  __nv_bfloat16* xhat;          // (rows x nO) normalised activations for the backward pass
  float* rstd;                  // (rows)
  float4* stats;                // (rows_pad, n_tiles) tagged partial sums {sum, sumsq, tag, 0}; zero-initialised once
  unsigned int* seq;            // [0] launch tag (starts at 1), [1] finished-CTA count, [2] time-out flag
  float drop_p;
  uint64_t seed;
  const int64_t* seed_dev;
};

struct GemmParams {
  int M, N, K;              // K = reduction length per shift (MODE_KK) or total (MODE_MNMN)
  int n_shifts;             // MODE_KK: 1 (plain) or 3 (window)
  int a_row_shift[3];       // added to the A row (M) coordinate
  int a_col_off[3];         // added to the A K coordinate
  int b_row_off[3];         // added to the B row coordinate (N for MODE_KK, K for MODE_KMN)
  int b_col_off[3];         // added to the B column coordinate (K for MODE_KK, N for MODE_KMN)
  int halo;                 // window GEMMs: A tile loaded ONCE per k-block with a one-row halo (see kernel)
  int splits;               // split-K factor (EPI_ATOMIC_F32)
  int win_w;                // MODE_MNMN: >0 = B is the window-expanded view of a (T, win_w) array
  const int* m_dev;         // optional device-side row count (<= M) for fixed-shape CUDA graphs
  void* out;                // bf16 (STORE / MAXOUT3) or fp32 (ATOMIC)
  int ldo;
  const __nv_bfloat16* bias;        // (N) or null
  uint8_t* which;                   // MAXOUT3: argmax piece per unit
  const __nv_bfloat16* add_src;     // STORE: optional out += row_scale[row] * add_src[row, n]
  const float* row_scale;
  int ld_add;
  // C2: optional consumer-side gate on the "published" flags of the buckets that hold B (the weights)
  // and the bias: the TMA producer warp waits for them right before its first load (gate.cuh)
  GateArgs gate;
  LnArgs ln;                        // EPI_MAXOUT3_LN only
};

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer);
// cluster: 1; 2 = pairs of CTAs on adjacent M tiles sharing the B operand by TMA multicast;
// 3 = the same pairs issuing one tcgen05.mma.cta_group::2 (M = 256), each CTA holding half of B.
// (For 2 and 3 the K-major B tensor map must be encoded with box rows block_n / 2.)
cudaError_t launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int block_n, int mode, int epi,
                        int cluster, int num_sms, cudaStream_t s);
bool gemm_supports_cluster(int block_n, int mode, int epi);
// window GEMM with shifts {-1,0,+1}: true if a halo variant (A loaded once per k-block) is compiled
bool gemm_supports_halo(int block_n, int mode, int epi, int cluster);
constexpr int kHaloRows = 130;   // A tensor map box rows for the halo variants
int gemm_block_k();

}  // namespace srb
