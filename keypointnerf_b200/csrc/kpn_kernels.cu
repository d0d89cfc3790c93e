// Ray-march kernels, engine 1: fp32 SIMT ("parity anchor").
//
// Stages (reference src/model.py:1018-1096 for the eval branch of batch_render_pifu_nerf):
//   pack_nhwc      feature maps NCHW -> channel-last atlases (once per scene)
//   prep_*         derived camera constants on device (no host sync)
//   rays           pixel lattice -> ray direction, near/far incl. bbox clip
//   coarse_z       uniform depths
//   compact        per-sample validity (frustum + foreground in every view) -> compacted work list
//   shade_simt     per-sample gather + keypoint encoding + MLPs for tiles of 64 valid samples
//   composite      alpha compositing along the ray
//   importance     inverse-CDF resampling + merge with the coarse depths
#include <atomic>
#include "kpn_device.cuh"
#include "kpn_launch.h"

namespace kpn {

// ------------------------------------------------------------------------------------------------
// scene packing
// ------------------------------------------------------------------------------------------------
// in (V,C,H,W) fp32 -> out [V][H][W][Cp] fp32, Cp >= C zero padded.
__global__ void pack_nhwc_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int V, int C, int H, int W, int Cp) {
  long long n = (long long)V * H * W * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    long long t = i / Cp;
    int x = (int)(t % W); t /= W;
    int y = (int)(t % H);
    int v = (int)(t / H);
    out[i] = c < C ? in[(((long long)v * C + c) * H + y) * W + x] : 0.0f;
  }
}

__device__ void invert_n(double* a, double* inv, int n) {
  // Gauss-Jordan with partial pivoting on an n x n (n <= 4) row-major matrix.
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(a[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); piv = r; }
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t;
        t = inv[c * n + j]; inv[c * n + j] = inv[piv * n + j]; inv[piv * n + j] = t;
      }
    double d = 1.0 / a[c * n + c];
    for (int j = 0; j < n; ++j) { a[c * n + j] *= d; inv[c * n + j] *= d; }
    for (int r = 0; r < n; ++r)
      if (r != c) {
        double f = a[r * n + c];
        for (int j = 0; j < n; ++j) { a[r * n + j] -= f * a[c * n + j]; inv[r * n + j] -= f * inv[c * n + j]; }
      }
  }
}

__global__ void prep_scene_kernel(const RawScene* __restrict__ raw, DevScene* sc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int v = 0; v < sc->V; ++v) {
    for (int i = 0; i < 12; ++i) { sc->P[v][i] = raw->KRT[v * 16 + i]; sc->E[v][i] = raw->extrin[v * 16 + i]; }
    double a[16], inv[16];
    for (int i = 0; i < 16; ++i) a[i] = (double)raw->KRT[v * 16 + i];
    invert_n(a, inv, 4);
    for (int i = 0; i < 3; ++i) sc->C[v][i] = (float)inv[i * 4 + 3];
    for (int k = 0; k < sc->K; ++k) {
      const float* p = raw->kpt3d + 3 * k;
      const float* E = sc->E[v];
      for (int i = 0; i < 3; ++i) sc->kc[v][k][i] = E[4 * i] * p[0] + E[4 * i + 1] * p[1] + E[4 * i + 2] * p[2] + E[4 * i + 3];
    }
  }
  for (int i = 0; i < 3; ++i) { sc->bounds[i] = raw->bounds[i] - 0.01f; sc->bounds[3 + i] = raw->bounds[3 + i] + 0.01f; }
}

__global__ void prep_target_kernel(const RawTarget* __restrict__ raw, DevTarget* tg) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a[9], inv[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i * 3 + j] = (double)raw->K[i * 4 + j];
  invert_n(a, inv, 3);
  for (int i = 0; i < 9; ++i) tg->invK[i] = (float)inv[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tg->R[i * 3 + j] = raw->RT[i * 4 + j];
  for (int j = 0; j < 3; ++j) {
    float s = 0.0f;
    for (int i = 0; i < 3; ++i) s += raw->RT[i * 4 + 3] * raw->RT[i * 4 + j];
    tg->o[j] = -s;
  }
}

// ------------------------------------------------------------------------------------------------
// rays (reference src/model.py:1018-1043, 1178-1237)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ray_for_pixel(const DevScene& sc, const DevTarget& tg, float px, float py,
                                              float d[3], float& n_r, float& f_r) {
  const float* iK = tg.invK;
  float dc[3], gn[3], gf[3];
  for (int i = 0; i < 3; ++i) {
    dc[i] = iK[i * 3 + 0] * px + iK[i * 3 + 1] * py + iK[i * 3 + 2];
    gn[i] = iK[i * 3 + 0] * (tg.znear * px) + iK[i * 3 + 1] * (tg.znear * py) + iK[i * 3 + 2] * tg.znear;
    gf[i] = iK[i * 3 + 0] * (tg.zfar * px) + iK[i * 3 + 1] * (tg.zfar * py) + iK[i * 3 + 2] * tg.zfar;
  }
  n_r = sqrtf(gn[0] * gn[0] + gn[1] * gn[1] + gn[2] * gn[2]);
  f_r = sqrtf(gf[0] * gf[0] + gf[1] * gf[1] + gf[2] * gf[2]);
  float w[3];
  for (int j = 0; j < 3; ++j) w[j] = dc[0] * tg.R[j] + dc[1] * tg.R[3 + j] + dc[2] * tg.R[6 + j];
  float nn = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 1e-12f);
  d[0] = w[0] / nn; d[1] = w[1] / nn; d[2] = w[2] / nn;
  // bbox slab test with the "exactly two face hits" rule
  float dd[3];
  for (int i = 0; i < 3; ++i) dd[i] = fabsf(d[i]) < 1e-5f ? 1e-5f : d[i];
  float ddn = sqrtf(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
  const float* b = sc.bounds;
  const float eps = 1e-6f;
  int cnt = 0;
  float tn = 3.0e38f, tf = -3.0e38f;
  for (int f = 0; f < 6; ++f) {
    int ax = f % 3;
    float t = (b[f] - tg.o[ax]) / dd[ax];
    float p[3] = {t * dd[0] + tg.o[0], t * dd[1] + tg.o[1], t * dd[2] + tg.o[2]};
    bool in = p[0] >= b[0] - eps && p[0] <= b[3] + eps && p[1] >= b[1] - eps && p[1] <= b[4] + eps &&
              p[2] >= b[2] - eps && p[2] <= b[5] + eps;
    if (in) {
      float e[3] = {p[0] - tg.o[0], p[1] - tg.o[1], p[2] - tg.o[2]};
      float dist = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) / ddn;
      tn = fminf(tn, dist); tf = fmaxf(tf, dist);
      ++cnt;
    }
  }
  bool hit = cnt == 2;
  if (hit && tn > n_r) n_r = tn;
  if (hit && tf < f_r) f_r = tf;
}

__global__ void rays_kernel(const DevScene* __restrict__ sc, const DevTarget* __restrict__ tg, int r0, int nr,
                            float* __restrict__ ray_d, float* __restrict__ ray_nf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr) return;
  int r = r0 + i;
  int ix = r % tg->nx, iy = r / tg->nx;
  float px = (float)(tg->x0 + tg->step * ix), py = (float)(tg->y0 + tg->step_y * iy);
  float d[3], n, f;
  ray_for_pixel(*sc, *tg, px, py, d, n, f);
  ray_d[3 * i + 0] = d[0]; ray_d[3 * i + 1] = d[1]; ray_d[3 * i + 2] = d[2];
  ray_nf[2 * i + 0] = n; ray_nf[2 * i + 1] = f;
}

__device__ __forceinline__ float linspace01(int i, int S) {
  // torch.linspace(0, 1, S) in fp32: symmetric evaluation from both ends.
  if (S <= 1) return 0.0f;
  float step = 1.0f / (float)(S - 1);
  return i < S / 2 ? step * (float)i : 1.0f - step * (float)(S - 1 - i);
}

__global__ void coarse_z_kernel(const float* __restrict__ ray_nf, int nr, int S, float* __restrict__ z) {
  long long n = (long long)nr * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / S), s = (int)(i % S);
    float a = ray_nf[2 * r], b = ray_nf[2 * r + 1];
    z[i] = a + (b - a) * linspace01(s, S);  // reference src/model.py:1045-1055
  }
}

// ------------------------------------------------------------------------------------------------
// validity + compaction
// ------------------------------------------------------------------------------------------------
// Early-ray termination (ert.eps > 0, ray mode): only the samples s of a ray with ert.s_lo <= s < ert.s_hi are looked at (the
// others keep what an earlier segment wrote), and a ray whose transmittance behind the earlier segments, 1 - ert.ray_alpha[ray],
// is below ert.eps gets alpha = 0 entries instead of being shaded: what it could still add to the pixel is < eps per channel.
__global__ void compact_kernel(const DevScene* __restrict__ sc, SampleSrc src, long long n, int query_mode,
                               int* __restrict__ list, int* __restrict__ counter, float* __restrict__ out5,
                               uint8_t* __restrict__ valid_out, ErtSegment ert) {
  const DevScene& S = *sc;
  __shared__ int s_cnt[8];
  __shared__ int s_base;
  for (long long base = (blockIdx.x * (long long)blockDim.x) ; base < n; base += (long long)gridDim.x * blockDim.x) {   // block-uniform trip count
    long long i = base + threadIdx.x;
    bool ok = false;
    bool in_seg = i < n;
    if (in_seg && ert.s_hi > 0) {
      const int sidx = (int)(i % src.S);
      in_seg = sidx >= ert.s_lo && sidx < ert.s_hi;
    }
    if (in_seg) {
      float p[3], d[3];
      Proj q[MAXV];
      fetch_sample(src, i, p, d);
      ok = sample_valid(S, p, q);
      if (ok && ert.ray_alpha != nullptr && 1.0f - ert.ray_alpha[i / src.S] < ert.eps) ok = false;   // ray already opaque
      if (!ok) {
        float* o = out5 + 5 * i;
        if (query_mode) { o[0] = 0.f; o[1] = 0.f; }
        else { o[0] = 0.f; o[1] = S.sdf_invalid; }  // alpha = 0, sdf = 0.1/nml_scale (src/model.py:982,996)
        o[2] = 0.f; o[3] = 0.f; o[4] = 0.f;
      }
      if (valid_out) valid_out[i] = ok ? 1 : 0;
    }
    // block-aggregated append: one atomic per 256 samples instead of one per warp (the single counter serialises in L2)
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_cnt[wid] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) { int c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
      s_base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    if (ok) list[s_base + s_cnt[wid] + __popc(m & ((1u << lane) - 1u))] = (int)i;
    __syncthreads();   // s_cnt / s_base are rewritten by the next iteration
  }
}

// ------------------------------------------------------------------------------------------------
// shade (fp32 SIMT): tiles of 64 valid samples, 256 threads
// ------------------------------------------------------------------------------------------------
constexpr int TS = 64;
constexpr int NT = 256;
constexpr int LDA = 232;  // 7*MAXK(=24 supported here)+64, multiple of 4
constexpr int LDB = 128;

enum { ACT_NONE = 0, ACT_SP = 1, ACT_ELU = 2 };

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == ACT_SP) return softplus100(x);
  if (ACT == ACT_ELU) return elu1(x);
  return x;
}

// out[r][n] = act(bias[n] + sum_k in[r][k] * Wt[k][n]) for the 64 rows of the tile.
// Warp w owns rows 8w..8w+7 (A operand is a warp-wide broadcast), lane owns columns lane+32j.
template <int ACT, int NJ>
__device__ __forceinline__ void dense_tile(const DevWeightsF32& W, int L, const float* sIn, int ldi, float* sOut, int ldo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = W.K[L], N = W.N[L];
  constexpr int ldw = NJ * 32;
  const float* __restrict__ wt = W.wt[L];
  float acc[8][NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    float b = (lane + 32 * j < N) ? __ldg(W.bias[L] + lane + 32 * j) : 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r][j] = b;
  }
  const float* a0 = sIn + warp * 8 * ldi;
  int k = 0;
  for (; k + 4 <= K; k += 4) {
    float4 a[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(a0 + r * ldi + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float w[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) w[j] = __ldg(wt + (k + kk) * ldw + lane + 32 * j);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[r][j] = fmaf(av, w[j], acc[r][j]);
      }
    }
  }
  for (; k < K; ++k) {
    float w[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) w[j] = __ldg(wt + k * ldw + lane + 32 * j);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float av = a0[r * ldi + k];
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[r][j] = fmaf(av, w[j], acc[r][j]);
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    int col = lane + 32 * j;
    if (col < N) {
#pragma unroll
      for (int r = 0; r < 8; ++r) sOut[(warp * 8 + r) * ldo + col] = apply_act<ACT>(acc[r][j]);
    }
  }
}

struct ShadeSmem {
  float* A;     // [TS][LDA]
  float* B;     // [TS][LDB]
  float* XV;    // [3][TS][64]   (later scratch S1,S2,S3)
  float* F;     // [3][TS][36]
  float* MV;    // [TS][72]
  float* Lat;   // [TS][24]
  float* RD;    // [3][TS][4]
  float* RGB;   // [3][TS][4]
  float* UV;    // [3][TS][4]  u, v, zn, boundary weight
  float* PW;    // [3][TS]
  float* Om;    // [3][TS]
  float* P;     // [TS][8]  p xyz, d xyz
  float* F8;    // [TS][8]
  float* Geo;   // [TS][2]
  float* Vis;   // [TS]
  float* Logit; // [3][TS]
  int* Id;      // [TS]
};

constexpr int SHADE_SMEM_FLOATS = TS * LDA + TS * LDB + 3 * TS * 64 + 3 * TS * 36 + TS * 72 + TS * 24 + 3 * TS * 4 * 3 +
                                  3 * TS * 2 + TS * 8 * 2 + TS * 2 + TS + 3 * TS + TS;
constexpr size_t SHADE_SMEM_BYTES = SHADE_SMEM_FLOATS * sizeof(float);

__global__ void __launch_bounds__(NT, 1)
shade_simt_kernel(const DevScene* __restrict__ scp, const DevWeightsF32* __restrict__ Wp, SampleSrc src,
                  const int* __restrict__ list, const int* __restrict__ count_ptr, int query_mode,
                  float* __restrict__ out5) {
  extern __shared__ float4 smem4[];
  float* sm = reinterpret_cast<float*>(smem4);
  ShadeSmem s;
  s.A = sm; sm += TS * LDA;
  s.B = sm; sm += TS * LDB;
  s.XV = sm; sm += 3 * TS * 64;
  s.F = sm; sm += 3 * TS * 36;
  s.MV = sm; sm += TS * 72;
  s.Lat = sm; sm += TS * 24;
  s.RD = sm; sm += 3 * TS * 4;
  s.RGB = sm; sm += 3 * TS * 4;
  s.UV = sm; sm += 3 * TS * 4;
  s.PW = sm; sm += 3 * TS;
  s.Om = sm; sm += 3 * TS;
  s.P = sm; sm += TS * 8;
  s.F8 = sm; sm += TS * 8;
  s.Geo = sm; sm += TS * 2;
  s.Vis = sm; sm += TS;
  s.Logit = sm; sm += 3 * TS;
  s.Id = reinterpret_cast<int*>(sm);
  float* S1 = s.XV;
  float* S2 = s.XV + TS * 64;
  float* S3 = s.XV + 2 * TS * 64;

  const DevScene& sc = *scp;
  const DevWeightsF32& W = *Wp;
  const int t = threadIdx.x;
  const int V = sc.V, K = sc.K;
  const int encd = (1 + 2 * sc.sp_level) * K;
  const int count = *count_ptr;

  for (int tile = blockIdx.x; tile * TS < count; tile += gridDim.x) {
    const int base = tile * TS;
    const int nrows = min(TS, count - base);
    // ---- 1. fetch the tile's samples (rows past the end replicate the last one)
    if (t < TS) {
      int id = list[base + min(t, nrows - 1)];
      s.Id[t] = id;
      float p[3], d[3];
      fetch_sample(src, id, p, d);
      float* o = s.P + t * 8;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = d[0]; o[4] = d[1]; o[5] = d[2];
    }
    __syncthreads();
    // ---- 2. projections and view weights (reference src/model.py:713-723,750-759; mask == 1 here)
    for (int idx = t; idx < V * TS; idx += NT) {
      int v = idx / TS, r = idx % TS;
      Proj q = project_view(sc, v, s.P + r * 8);
      float* o = s.UV + idx * 4;
      o[0] = q.u; o[1] = q.v; o[2] = q.zn; o[3] = boundary_weight(q);
    }
    __syncthreads();
    if (t < TS) {
      float sum = 0.0f;
      for (int v = 0; v < V; ++v) sum += s.UV[(v * TS + t) * 4 + 3];
      for (int v = 0; v < V; ++v) s.PW[v * TS + t] = s.UV[(v * TS + t) * 4 + 3] / (sum + 1e-6f);
    }
    // ---- 3. per-view geometry MLP (reference src/utils.py:691-720)
    for (int v = 0; v < V; ++v) {
      {
        const int r = t >> 2, q4 = t & 3;
        const float* uv = s.UV + (v * TS + r) * 4;
        float* row = s.A + r * LDA;
        Taps t64 = make_taps(uv[0], uv[1], sc.f64.W, sc.f64.H);
        float f[16];
        gather_f32<4>(sc.f64, v, t64, q4 * 4, f);
#pragma unroll
        for (int i = 0; i < 16; ++i) row[encd + q4 * 16 + i] = f[i];
        float c[3];
        to_camera(sc, v, s.P + r * 8, c);
        for (int k = q4; k < K; k += 4) encode_kpt(sc, v, k, c, row);
        if (q4 < 2) {
          Taps t8 = make_taps(uv[0], uv[1], sc.f8.W, sc.f8.H);
          float g[4];
          gather_f32<1>(sc.f8, v, t8, q4, g);
#pragma unroll
          for (int i = 0; i < 4; ++i) s.F8[r * 8 + q4 * 4 + i] = g[i];
        }
      }
      __syncthreads();
      dense_tile<ACT_SP, 4>(W, L_GEO0, s.A, LDA, s.B, LDB);
      __syncthreads();
      dense_tile<ACT_SP, 4>(W, L_GEO1, s.B, LDB, s.A, LDA);
      for (int idx = t; idx < TS * 8; idx += NT) s.A[(idx >> 3) * LDA + 128 + (idx & 7)] = s.F8[idx];
      __syncthreads();
      dense_tile<ACT_SP, 4>(W, L_GEO2, s.A, LDA, s.B, LDB);
      __syncthreads();
      dense_tile<ACT_NONE, 2>(W, L_GEO3, s.B, LDB, s.XV + v * TS * 64, 64);
      __syncthreads();
    }
    // ---- 4. weighted mean || variance over views (reference src/utils.py:722-748), density head
    {
      const int r = t >> 2, q4 = t & 3;
      for (int c = q4 * 16; c < q4 * 16 + 16; ++c) {
        float mean = 0.0f;
        for (int v = 0; v < V; ++v) mean += s.PW[v * TS + r] * s.XV[(v * TS + r) * 64 + c];
        float var = 0.0f;
        for (int v = 0; v < V; ++v) {
          float dlt = s.XV[(v * TS + r) * 64 + c] - mean;
          var += s.PW[v * TS + r] * dlt * dlt;
        }
        s.A[r * LDA + c] = mean;
        s.A[r * LDA + 64 + c] = var;
      }
    }
    __syncthreads();
    dense_tile<ACT_NONE, 1>(W, L_CMP, s.A, LDA, s.Lat, 24);   // ibr_compress_gfeat (src/model.py:819)
    dense_tile<ACT_SP, 2>(W, L_DEN0, s.A, LDA, s.B, LDB);
    __syncthreads();
    dense_tile<ACT_SP, 2>(W, L_DEN1, s.B, LDB, s.A, LDA);
    __syncthreads();
    dense_tile<ACT_NONE, 1>(W, L_DEN2, s.A, LDA, s.Geo, 2);
    // ---- 5. colour branch inputs (reference src/model.py:806-832)
    for (int idx = t; idx < V * TS; idx += NT) {
      int v = idx / TS, r = idx % TS;
      const float* uv = s.UV + idx * 4;
      float* f = s.F + idx * 36;
      Taps ti = make_taps(uv[0], uv[1], sc.img.W, sc.img.H);
      float c4[4];
      gather_f32<1>(sc.img, v, ti, 0, c4);
      s.RGB[idx * 4 + 0] = c4[0]; s.RGB[idx * 4 + 1] = c4[1]; s.RGB[idx * 4 + 2] = c4[2];
      f[0] = c4[0]; f[1] = c4[1]; f[2] = c4[2];
      Taps tt = make_taps(uv[0], uv[1], sc.ftex.W, sc.ftex.H);
      float g[8];
      gather_f32<2>(sc.ftex, v, tt, 0, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[3 + i] = g[i];
      ray_diff(sc, v, s.P + r * 8, s.P + r * 8 + 3, s.RD + idx * 4);
    }
    __syncthreads();
    for (int idx = t; idx < V * TS * 24; idx += NT) {
      int c = idx % 24, vr = idx / 24;
      s.F[vr * 36 + 11 + c] = s.Lat[(vr % TS) * 24 + c];
    }
    // ray-direction encoder 4->16->35 (ELU) added onto the features (src/model.py:1279-1284)
    for (int v = 0; v < V; ++v) {
      __syncthreads();
      dense_tile<ACT_ELU, 1>(W, L_RE0, s.RD + v * TS * 4, 4, s.B, LDB);
      __syncthreads();
      dense_tile<ACT_ELU, 2>(W, L_RE1, s.B, LDB, s.A, LDA);
      __syncthreads();
      for (int idx = t; idx < TS * 35; idx += NT) {
        int r = idx / 35, c = idx % 35;
        s.F[(v * TS + r) * 36 + c] += s.A[r * LDA + c];
      }
    }
    // blending weights (src/model.py:1286-1289), mask == 1
    if (t < TS) {
      float ex[MAXV], mn = 3.0e38f;
      for (int v = 0; v < V; ++v) {
        ex[v] = expf(W.ani_al_abs * (s.RD[(v * TS + t) * 4 + 3] - 1.0f));
        mn = fminf(mn, ex[v]);
      }
      float sum = 0.0f;
      for (int v = 0; v < V; ++v) { ex[v] -= mn; sum += ex[v]; }
      for (int v = 0; v < V; ++v) s.Om[v * TS + t] = ex[v] / (sum + 1e-8f);
    }
    __syncthreads();
    for (int idx = t; idx < TS * 35; idx += NT) {
      int r = idx / 35, c = idx % 35;
      float mean = 0.0f;
      for (int v = 0; v < V; ++v) mean += s.F[(v * TS + r) * 36 + c] * s.Om[v * TS + r];
      float var = 0.0f;
      for (int v = 0; v < V; ++v) {
        float dlt = s.F[(v * TS + r) * 36 + c] - mean;
        var += s.Om[v * TS + r] * dlt * dlt;
      }
      s.MV[r * 72 + c] = mean;
      s.MV[r * 72 + 35 + c] = var;
    }
    __syncthreads();
    // ---- 6. per-view IBR head (src/model.py:1292-1300)
    for (int v = 0; v < V; ++v) {
      for (int idx = t; idx < TS * 105; idx += NT) {
        int r = idx / 105, c = idx % 105;
        s.A[r * LDA + c] = c < 70 ? s.MV[r * 72 + c] : s.F[(v * TS + r) * 36 + (c - 70)];
      }
      __syncthreads();
      dense_tile<ACT_ELU, 2>(W, L_BASE0, s.A, LDA, s.B, LDB);
      __syncthreads();
      dense_tile<ACT_ELU, 1>(W, L_BASE1, s.B, LDB, S1, 64);
      __syncthreads();
      for (int idx = t; idx < TS * 32; idx += NT) {
        int r = idx >> 5, c = idx & 31;
        S2[r * 64 + c] = S1[r * 64 + c] * s.Om[v * TS + r];
      }
      __syncthreads();
      dense_tile<ACT_ELU, 1>(W, L_VIS1A, S2, 64, S3, 64);
      __syncthreads();
      dense_tile<ACT_ELU, 2>(W, L_VIS1B, S3, 64, S2, 64);
      __syncthreads();
      for (int idx = t; idx < TS * 32; idx += NT) {
        int r = idx >> 5, c = idx & 31;
        float x = S1[r * 64 + c] + S2[r * 64 + c];
        S1[r * 64 + c] = x;
        S3[r * 64 + c] = x * sigmoidf(S2[r * 64 + 32]);
      }
      __syncthreads();
      dense_tile<ACT_ELU, 1>(W, L_VIS2A, S3, 64, S2, 64);
      __syncthreads();
      dense_tile<ACT_NONE, 1>(W, L_VIS2B, S2, 64, s.Vis, 1);
      __syncthreads();
      for (int idx = t; idx < TS * 37; idx += NT) {
        int r = idx / 37, c = idx % 37;
        float val = c < 32 ? S1[r * 64 + c] : (c == 32 ? sigmoidf(s.Vis[r]) : s.RD[(v * TS + r) * 4 + (c - 33)]);
        S3[r * 64 + c] = val;
      }
      __syncthreads();
      dense_tile<ACT_ELU, 1>(W, L_OUT0, S3, 64, S2, 64);
      __syncthreads();
      dense_tile<ACT_ELU, 1>(W, L_OUT1, S2, 64, S3, 64);
      __syncthreads();
      dense_tile<ACT_NONE, 1>(W, L_OUT2, S3, 64, s.Logit + v * TS, 1);
      __syncthreads();
    }
    // ---- 7. softmax blend of the source colours and output (src/model.py:1300-1301, 978-997)
    if (t < nrows) {
      float m = -3.0e38f;
      for (int v = 0; v < V; ++v) m = fmaxf(m, s.Logit[v * TS + t]);
      float den = 0.0f, rgb[3] = {0.f, 0.f, 0.f};
      for (int v = 0; v < V; ++v) {
        float e = expf(s.Logit[v * TS + t] - m);
        den += e;
        for (int c = 0; c < 3; ++c) rgb[c] += e * s.RGB[(v * TS + t) * 4 + c];
      }
      float g0 = s.Geo[t * 2], rad = s.Geo[t * 2 + 1];
      float* o = out5 + 5ll * s.Id[t];
      if (query_mode) { o[0] = g0; o[1] = rad; }
      else { o[0] = fmaxf(rad, 0.0f); o[1] = g0; }
      o[2] = rgb[0] / den; o[3] = rgb[1] / den; o[4] = rgb[2] / den;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// compositing (reference src/model.py:1150-1176)
// ------------------------------------------------------------------------------------------------
// rgba (nr,S,5) = [alpha, sdf, r, g, b]; z (nr,S).  Planar outputs indexed by global ray r0+i.
// One WARP per ray: lane l takes samples l, l+32, ... (coalesced reads), the transmittance T_k = prod_{j<k} (1 - a_j) is an
// exclusive product scan across the lanes carried from one group of 32 samples to the next, the ray sums are warp reductions.
__global__ void __launch_bounds__(128)
composite_kernel(const float* __restrict__ rgba, const float* __restrict__ z, int r0, int nr, int S, int S_eval,
                 long long plane, float* __restrict__ color, float* __restrict__ depth,
                 float* __restrict__ alpha, float* __restrict__ sdf, float* __restrict__ contrib) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nr; i += gridDim.x * wpb) {
    const float* q = rgba + (long long)i * S * 5;
    const float* zz = z + (long long)i * S;
    float Tc = 1.0f;   // transmittance in front of this group of 32 samples
    float acc = 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, cd = 0.0f, cs = 0.0f;
    for (int k0 = 0; k0 < S_eval; k0 += 32) {   // S_eval < S: composite of the first S_eval samples only (ERT segment)
      const int k = k0 + lane;
      float a = 0.0f, zk = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f, q4 = 0.0f;
      if (k < S_eval) {
        zk = zz[k];
        const float dist = k + 1 < S ? zz[k + 1] - zk : 1e10f;
        a = 1.0f - expf(-q[5 * k] * dist);
        q1 = q[5 * k + 1]; q2 = q[5 * k + 2]; q3 = q[5 * k + 3]; q4 = q[5 * k + 4];
      }
      float p = 1.0f - a;   // inclusive product scan of (1 - a)
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, p, d);
        if (lane >= d) p *= t;
      }
      float ex = __shfl_up_sync(0xffffffffu, p, 1);
      if (lane == 0) ex = 1.0f;
      const float c = a * (Tc * ex);
      Tc *= __shfl_sync(0xffffffffu, p, 31);
      if (k < S_eval) {
        acc += c; cr += c * q2; cg += c * q3; cb += c * q4; cd += c * zk; cs += c * q1;
        if (contrib) contrib[(long long)i * S + k] = c;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, d); cr += __shfl_xor_sync(0xffffffffu, cr, d); cg += __shfl_xor_sync(0xffffffffu, cg, d);
      cb += __shfl_xor_sync(0xffffffffu, cb, d); cd += __shfl_xor_sync(0xffffffffu, cd, d); cs += __shfl_xor_sync(0xffffffffu, cs, d);
    }
    if (lane == 0) {
      const long long r = r0 + i;
      if (color) { color[r] = cr; color[plane + r] = cg; color[2 * plane + r] = cb; }
      if (alpha) alpha[r] = acc;
      if (depth) depth[r] = cd / (acc + 1e-8f);
      if (sdf) sdf[r] = cs / (acc + 1e-8f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hierarchical resampling (reference src/model.py:1110-1148, 1072-1076), uniform=True
// ------------------------------------------------------------------------------------------------
constexpr int MAX_SC = 256;

__global__ void importance_kernel(const float* __restrict__ contrib, const float* __restrict__ z, int nr, int Sc, int Sf,
                                  float* __restrict__ zout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr) return;
  const float* c = contrib + (long long)i * Sc;
  const float* zz = z + (long long)i * Sc;
  float* out = zout + (long long)i * (Sc + Sf);
  const int nb = Sc - 2;   // pdf bins = contrib[1:-1]
  const int nc = Sc - 1;   // cdf entries == z_mid entries
  float cdf[MAX_SC];
  float sum = 0.0f;
  for (int k = 0; k < nb; ++k) sum += c[k + 1] + 1e-5f;
  cdf[0] = 0.0f;
  float run = 0.0f;
  for (int k = 0; k < nb; ++k) { run += (c[k + 1] + 1e-5f) / sum; cdf[k + 1] = run; }
  // fine depths are written after the coarse ones, then merged
  int idx = 0;  // number of cdf entries <= u (searchsorted right=True); u is increasing so idx only grows
  for (int j = 0; j < Sf; ++j) {
    float u = linspace01(j, Sf);
    while (idx < nc && cdf[idx] <= u) ++idx;
    int lo = max(idx - 1, 0), hi = min(idx, nc - 1);
    float clo = cdf[lo], chi = cdf[hi];
    float zlo = 0.5f * (zz[lo + 1] + zz[lo]), zhi = 0.5f * (zz[hi + 1] + zz[hi]);
    float den = chi - clo;
    if (den < 1e-5f) den = 1.0f;
    out[Sc + j] = zlo + ((u - clo) / den) * (zhi - zlo);
  }
  for (int k = 0; k < Sc; ++k) out[k] = zz[k];
  // sort(cat[z, z_fine]): insertion sort (both halves are already ordered, so this is a merge)
  for (int a = Sc; a < Sc + Sf; ++a) {
    float val = out[a];
    int b = a - 1;
    while (b >= 0 && out[b] > val) { out[b + 1] = out[b]; --b; }
    out[b + 1] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline int grid_for(long long n, int block, int cap) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

cudaError_t launch_pack_nhwc_f32(const float* in, float* out, int V, int C, int H, int W, int Cp, cudaStream_t st) {
  long long n = (long long)V * H * W * Cp;
  pack_nhwc_f32_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(in, out, V, C, H, W, Cp);
  return cudaGetLastError();
}
cudaError_t launch_prep_scene(const RawScene* raw, DevScene* sc, cudaStream_t st) {
  prep_scene_kernel<<<1, 32, 0, st>>>(raw, sc);
  return cudaGetLastError();
}
cudaError_t launch_prep_target(const RawTarget* raw, DevTarget* tg, cudaStream_t st) {
  prep_target_kernel<<<1, 32, 0, st>>>(raw, tg);
  return cudaGetLastError();
}
cudaError_t launch_rays(const DevScene* sc, const DevTarget* tg, int r0, int nr, float* ray_d, float* ray_nf, cudaStream_t st) {
  rays_kernel<<<grid_for(nr, 128, 1 << 30), 128, 0, st>>>(sc, tg, r0, nr, ray_d, ray_nf);
  return cudaGetLastError();
}
cudaError_t launch_coarse_z(const float* ray_nf, int nr, int S, float* z, cudaStream_t st) {
  coarse_z_kernel<<<grid_for((long long)nr * S, 256, 148 * 16), 256, 0, st>>>(ray_nf, nr, S, z);
  return cudaGetLastError();
}
cudaError_t launch_compact(const DevScene* sc, const SampleSrc& src, long long n, int query_mode, int* list, int* counter,
                           float* out5, uint8_t* valid_out, const ErtSegment& ert, cudaStream_t st) {
  compact_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(sc, src, n, query_mode, list, counter, out5, valid_out, ert);
  return cudaGetLastError();
}
cudaError_t launch_shade_simt(const DevScene* sc, const DevWeightsF32* W, const SampleSrc& src, const int* list,
                              const int* counter, long long n_max, int query_mode, float* out5, int num_sms,
                              cudaStream_t st) {
  static std::atomic<bool> attr_set[64];   // function attributes are per device (setting them twice is harmless)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(shade_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SHADE_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  int grid = grid_for(n_max, TS, num_sms);
  shade_simt_kernel<<<grid, NT, SHADE_SMEM_BYTES, st>>>(sc, W, src, list, counter, query_mode, out5);
  return cudaGetLastError();
}
cudaError_t launch_composite(const float* rgba, const float* z, int r0, int nr, int S, int S_eval, long long plane, float* color,
                             float* depth, float* alpha, float* sdf, float* contrib, cudaStream_t st) {
  composite_kernel<<<grid_for((long long)nr * 32, 128, 148 * 64), 128, 0, st>>>(rgba, z, r0, nr, S, S_eval, plane, color, depth, alpha, sdf, contrib);
  return cudaGetLastError();
}
cudaError_t launch_importance(const float* contrib, const float* z, int nr, int Sc, int Sf, float* zout, cudaStream_t st) {
  importance_kernel<<<grid_for(nr, 64, 1 << 30), 64, 0, st>>>(contrib, z, nr, Sc, Sf, zout);
  return cudaGetLastError();
}
int max_coarse_samples() { return MAX_SC; }
int simt_max_kpt() { return (LDA - 64) / 7; }

}  // namespace kpn
