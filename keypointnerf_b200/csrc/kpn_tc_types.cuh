// Tensor-core engine: stage table (which dense layers run on tcgen05, their padded shapes and where their
// fp16 weights sit in the shared-memory blob) and the fp32 constants that travel as kernel parameters
// (biases of the tensor-core layers + the tiny layers that stay on the CUDA cores in fp32).
#pragma once
#include <stdint.h>

namespace kpn {

constexpr int TC_NSTAGE = 12;
// stage:            0 L0   1 L1  2 L2  3 L3  4 P0|CMP 5 P1 6 BASE0 7 BASE1 8 VIS1A 9 VIS1B 10 VIS2A 11 OUT0
// Np (padded N):    128    128   128   64    96       64   64      32      32      48      32       16
// Kp (padded K):    K0P    128   144   128   128      64   112     64      32      32      32       48

struct TcStage {
  int Kp, Np;
  uint32_t off;  // byte offset of the stage's weight tile in the blob
};

struct TcPlan {
  TcStage st[TC_NSTAGE];
  uint32_t total_bytes;
  int K0P;  // padded input width of layer 0
};

__host__ __device__ constexpr int tc_k0p(int n_kpt) { return ((7 * n_kpt + 64) + 15) / 16 * 16; }

__host__ __device__ constexpr TcPlan make_tc_plan(int n_kpt) {
  TcPlan p{};
  const int Kp[TC_NSTAGE] = {tc_k0p(n_kpt), 128, 144, 128, 128, 64, 112, 64, 32, 32, 32, 48};
  const int Np[TC_NSTAGE] = {128, 128, 128, 64, 96, 64, 64, 32, 32, 48, 32, 16};
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
  float b_l0[128], b_l1[128], b_l2[128], b_l3[64];
  float b_p0[64], b_cmp[32], b_p1[64];
  float w_p2[2][64], b_p2[2];          // density head last layer (fp32 on CUDA cores)
  float w_re0[16][4], b_re0[16];       // ray-direction encoder
  float w_re1[35][16], b_re1[35];
  float b_base0[64], b_base1[32], b_vis1a[32], b_vis1b[48], b_vis2a[32];
  float w_vis2b[32], b_vis2b;
  float b_out0[16];
  float w_out1[8][16], b_out1[8];
  float w_out2[8], b_out2;
  float ani_abs;
};

}  // namespace kpn
