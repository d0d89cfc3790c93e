// Tensor-core engine: stage table (which dense layers run on tcgen05, their padded shapes and where their
// fp16 weights sit in the shared-memory blob) and the fp32 constants that travel as kernel parameters
// (biases of the tensor-core layers + the tiny layers that stay on the CUDA cores in fp32).
#pragma once
#include <stdint.h>

namespace kpn {

constexpr int TC_NSTAGE = 13;
// stage:            0 L0   1 L1  2 L2  3 L3  4 P0|CMP 5 P1 6 BASE0 7 BASE1 8 VIS1A 9 VIS1B 10 VIS2A 11 OUT0 12 RE1
// Np (padded N):    128    128   128   64    96       64   64      32      32      48      32       16      48
// Kp (padded K):    K0P    144   144   128   144      80   112     64      32      32      32       48      32
// Stage 12 (second ray-encoder layer 16->35 of the colour head, executed FIRST in the colour kernel) also carries its bias row.
// The geometry stages (0..5) carry their bias as one extra K row (the activation tile holds a constant 1 there), so their
// epilogues have no per-column constants.  Stage 0 additionally permutes its inputs (tc_kmap) so that the two threads
// that build a row's input each write one contiguous, naturally aligned run of tensor-memory columns.

struct TcStage {
  int Kp, Np;
  uint32_t off;  // byte offset of the stage's weight tile in the blob
};

struct TcPlan {
  TcStage st[TC_NSTAGE];
  uint32_t total_bytes;
  int K0P;  // padded input width of layer 0
};

// layer-0 input width incl. the bias row, padded to a multiple of 16
__host__ __device__ constexpr int tc_k0p(int n_kpt) { return ((7 * n_kpt + 64 + 1) + 15) / 16 * 16; }
// split of the layer-0 input between the two threads of a row: thread 0 builds keypoint pairs [0, PA) and feature
// channels [0, FA), thread 1 the remaining pairs, the remaining channels and the bias column
// (18 keypoints: 4 | 5 pairs; 24: 4 | 8 pairs -- the second thread starts a tile earlier, its partner writes the tile's outputs;
// the 40 | 24 channel split keeps one gather plan for both keypoint counts: 10 | 6 float4 groups of feat64)
__host__ __device__ constexpr int tc_l0_pa(int n_kpt) { return 4 + 0 * n_kpt; }
__host__ __device__ constexpr int tc_l0_fa(int n_kpt) { return 40 + 0 * n_kpt; }

// K index (fp16 element of the activation row) that input `i` of geometry stage `stage` is multiplied with.
// Stage 0 inputs: i < 7*n_kpt is encoding element r*n_kpt + k (reference src/spatial.py layout), else feat64 channel.
__host__ __device__ constexpr int tc_kmap(int stage, int n_kpt, int i) {
  if (stage != 0) return i;
  const int NP = n_kpt / 2, PA = tc_l0_pa(n_kpt), FA = tc_l0_fa(n_kpt);
  if (i < 7 * n_kpt) {
    const int r = i / n_kpt, k = i % n_kpt, j = k / 2;
    const int col = j < PA ? 7 * j + r : 7 * PA + FA / 2 + 7 * (j - PA) + r;
    return 2 * col + (k & 1);
  }
  const int c = i - 7 * n_kpt;
  return c < FA ? 2 * 7 * PA + c : 2 * (7 * PA + FA / 2 + 7 * (NP - PA)) + (c - FA);
}
// K index of the bias row of a geometry stage (or of stage 12)
__host__ __device__ constexpr int tc_kbias(int stage, int n_kpt) {
  switch (stage) {
    case 0: return 2 * (7 * (n_kpt / 2) + 32);
    case 1: return 128;
    case 2: return 136;
    case 3: return 120;
    case 4: return 128;
    case 12: return 16;
    default: return 64;
  }
}

// ---- "view-sequential" geometry kernel (18 keypoints; engines 0 and 3): a row is a SAMPLE, its three views are run through
// stages 0-3 one after the other, FOUR threads build / post-process a row (column quarters).  Layer-0 input of one
// (sample, view): 96 packed columns = 4 runs of 24: threads 0-2: two keypoint pairs (14 columns) + five float4 groups of feat64
// (10 columns); thread 3: three pairs (21) + one group (2) + the bias column.
__host__ __device__ constexpr int vs_run_cols(int n_kpt) { return n_kpt == 18 ? 24 : 0; }
__host__ __device__ constexpr int tc_kmap_vseq(int stage, int n_kpt, int i) {
  if (stage != 0) return i;
  if (i < 7 * n_kpt) {
    const int r = i / n_kpt, k = i % n_kpt, j = k / 2;
    const int t = j < 6 ? j / 2 : 3;
    const int col = 24 * t + (j < 6 ? (j % 2) * 7 : (j - 6) * 7) + r;
    return 2 * col + (k & 1);
  }
  const int c = i - 7 * n_kpt, g = c / 4;
  const int t = g < 15 ? g / 5 : 3, gi = g < 15 ? g % 5 : 0;
  const int col = 24 * t + (t < 3 ? 14 : 21) + 2 * gi + (c % 4) / 2;
  return 2 * col + (c & 1);
}
__host__ __device__ constexpr int tc_kbias_vseq(int stage, int n_kpt) { return stage == 0 ? 2 * 95 : tc_kbias(stage, n_kpt); }

__host__ __device__ constexpr TcPlan make_tc_plan(int n_kpt) {
  TcPlan p{};
  const int Kp[TC_NSTAGE] = {tc_k0p(n_kpt), 144, 144, 128, 144, 80, 112, 64, 32, 32, 32, 48, 32};
  const int Np[TC_NSTAGE] = {128, 128, 128, 64, 96, 64, 64, 32, 32, 48, 32, 16, 48};
  uint32_t off = 0;
  for (int i = 0; i < TC_NSTAGE; ++i) {
    p.st[i].Kp = Kp[i];
    p.st[i].Np = Np[i];
    p.st[i].off = off;
    off += (uint32_t)(Kp[i] * Np[i] * 2);
  }
  p.total_bytes = off;
  p.K0P = Kp[0];
  return p;
}

// fp32 constants, passed by value as a __grid_constant__ kernel parameter so that unrolled epilogues read
// them as constant-bank operands.
struct TcConsts {
  float w_p2[2][64], b_p2[2];          // density head last layer (fp32 on CUDA cores); the geometry-stage biases ride in the MMAs
  float w_re0[16][4], b_re0[16];       // ray-direction encoder, first layer (the second runs on the tensor core, stage 12)
  float b_base0[64], b_base1[32], b_vis1a[32], b_vis1b[48], b_vis2a[32];
  float w_vis2b[32], b_vis2b;
  float b_out0[16];
  float w_out1[8][16], b_out1[8];
  float w_out2[8], b_out2;
  float ani_abs;
};

}  // namespace kpn
