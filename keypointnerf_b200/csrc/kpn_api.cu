// C ABI of the ray-march path (include/kpnerf_b200.h): context, weight packer, scene packer,
// render / query drivers.  Host logic only; kernels live in kpn_kernels.cu (fp32 SIMT engine)
// and kpn_tc.cu (tensor-core engine).
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "kpn_launch.h"
#include "kpn_tc.cuh"

using namespace kpn;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int MAX_CHUNKS = 2048;

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct kpn_ctx {
  int device = 0;
  int num_sms = 148;
  std::string err;
  bool have_weights = false, have_scene = false;
  int n_kpt = 0;
  int sp_level = 3;
  float sp_scale = 1.0f, sp_sigma = 0.1f;

  RawScene* d_raw_scene = nullptr;
  DevScene* d_scene = nullptr;
  RawTarget* d_raw_target = nullptr;
  DevTarget* d_target = nullptr;
  DevWeightsF32* d_wf32 = nullptr;
  DevBuf wbuf;
  DevBuf wblob;                // fp16 weight tiles of the tensor-core engine (core-matrix layout)
  DevBuf wlo;                  // per-CTA-rank half-blobs [W_hi halves | W_lo halves] of the geometry stages (CTA-pair kernel)
  DevBuf wlo_vs;               // the same for the view-sequential kernel (different layer-0 input permutation)
  TcConsts tcc;                // fp32 constants of the tensor-core engine (kernel parameter)
  bool tc_weights = false;
  int scene_views = 0;
  int* d_counters = nullptr;   // [MAX_CHUNKS] valid samples per shading batch
  int* d_counters2 = nullptr;  // [MAX_CHUNKS] samples that needed a colour (tensor-core engine)
  int counters_used = 0;
  unsigned long long last_total = 0;

  DevBuf stage[4];             // host-sourced maps before packing
  DevBuf atlas[4];             // f64, f8, ftex, img (channel-last fp32)
  DevBuf atlas_fg;
  // per-chunk workspace: rays (dir, near/far), per-ray list ranges of up to two segments, the work list, the compact per-sample
  // records (alpha+sdf, rgb, compositing weight), the fine depths, the colour kernel's latent scratch + work list
  DevBuf ws_rayd, ws_raynf, ws_rayseg, ws_list, ws_ao, ws_rgb, ws_cw, ws_zfine, ws_lat, ws_list2, ws_ert, ws_contrib;
  DevBuf ws_out, ws_in;
  unsigned long long launches = 0;
  bool profiling = false;
  unsigned int* h_wd = nullptr;   // pinned host copy of the tensor-core kernels' watchdog words (refreshed by every render/query)
  std::vector<cudaEvent_t> ev_pool;   // pairs: [2i] start, [2i+1] stop
  size_t ev_used = 0;
};

#define KPN_FAIL(ctx, code, ...)                                   \
  do {                                                             \
    char _b[512];                                                  \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                         \
    (ctx)->err = _b;                                               \
    return (code);                                                 \
  } while (0)

#define KPN_CUDA(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) KPN_FAIL(ctx, KPN_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

extern "C" int kpn_abi_version(void) { return KPN_ABI_VERSION; }

extern "C" int kpn_create(int device, kpn_ctx** out) {
  if (!out) return KPN_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return KPN_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= ndev) return KPN_ERR_ARG;
  kpn_ctx* c = new kpn_ctx();
  c->device = device;
  DeviceGuard g(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return KPN_ERR_CUDA; }
  c->num_sms = prop.multiProcessorCount;
  bool ok = cudaMalloc(&c->d_raw_scene, sizeof(RawScene)) == cudaSuccess &&
            cudaMalloc(&c->d_scene, sizeof(DevScene)) == cudaSuccess &&
            cudaMalloc(&c->d_raw_target, sizeof(RawTarget)) == cudaSuccess &&
            cudaMalloc(&c->d_target, sizeof(DevTarget)) == cudaSuccess &&
            cudaMalloc(&c->d_wf32, sizeof(DevWeightsF32)) == cudaSuccess &&
            cudaMalloc(&c->d_counters, sizeof(int) * MAX_CHUNKS) == cudaSuccess &&
            cudaMalloc(&c->d_counters2, sizeof(int) * MAX_CHUNKS) == cudaSuccess;
  if (!ok) { kpn_destroy(c); return KPN_ERR_CUDA; }
  cudaMemset(c->d_counters, 0, sizeof(int) * MAX_CHUNKS);
  cudaMemset(c->d_counters2, 0, sizeof(int) * MAX_CHUNKS);
  if (cudaHostAlloc(reinterpret_cast<void**>(&c->h_wd), 8 * sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess) {
    kpn_destroy(c);
    return KPN_ERR_CUDA;
  }
  memset(c->h_wd, 0, 8 * sizeof(unsigned int));
  *out = c;
  return KPN_OK;
}

extern "C" void kpn_destroy(kpn_ctx* c) {
  if (!c) return;
  DeviceGuard g(c->device);
  cudaFree(c->d_raw_scene); cudaFree(c->d_scene); cudaFree(c->d_raw_target); cudaFree(c->d_target);
  cudaFree(c->d_wf32); cudaFree(c->d_counters); cudaFree(c->d_counters2);
  c->wbuf.release();
  c->wblob.release();
  c->wlo.release();
  c->wlo_vs.release();
  for (auto& b : c->stage) b.release();
  for (auto& b : c->atlas) b.release();
  c->atlas_fg.release();
  for (DevBuf* b : {&c->ws_rayd, &c->ws_raynf, &c->ws_rayseg, &c->ws_list, &c->ws_ao, &c->ws_rgb, &c->ws_cw, &c->ws_zfine, &c->ws_lat,
                    &c->ws_list2, &c->ws_ert, &c->ws_contrib, &c->ws_out, &c->ws_in})
    b->release();
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->h_wd) cudaFreeHost(c->h_wd);
  delete c;
}

extern "C" const char* kpn_last_error(const kpn_ctx* c) { return c ? c->err.c_str() : "null context"; }

// Device watchdog (kpn_tc.cuh): a barrier wait that gives up raises a sticky device flag; every later wait of every later
// tensor-core launch then aborts after a few polls, i.e. the engine keeps producing garbage until the flag is cleared.  Every
// render/query ends with an asynchronous copy of the flag words into pinned host memory (wd_snapshot); every entry point starts
// by looking at the last copy that has landed (wd_check): a raised flag is reported ONCE as KPN_ERR_CUDA and cleared on the
// device (stream-ordered), which re-arms the engine for the calls that follow.
static int wd_check(kpn_ctx* c, cudaStream_t st) {
  volatile unsigned int* w = c->h_wd;
  if (!w || w[0] == 0u) return KPN_OK;
  const unsigned int blk = w[1], thr = w[2], tag = w[3], par = w[4];
  for (int i = 0; i < 8; ++i) c->h_wd[i] = 0u;
  KPN_CUDA(c, tc_watchdog_clear_async(st));
  KPN_FAIL(c, KPN_ERR_CUDA, "device watchdog: a barrier wait of a tensor-core kernel gave up (block %u, thread %u, tag 0x%x, parity %u); "
           "the results of that call (and of tensor-core calls enqueued after it) are invalid; the flag has been cleared",
           blk, thr, tag, par);
}
static int wd_snapshot(kpn_ctx* c, cudaStream_t st) {
  KPN_CUDA(c, tc_watchdog_read_async(c->h_wd, st));
  return KPN_OK;
}

extern "C" int kpn_check_health(kpn_ctx* c, void* stream) {
  if (!c) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  return wd_check(c, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
static void expected_dims(int n_kpt, int sp_level, int dims[NLAYER][2]) {
  const int enc = (1 + 2 * sp_level) * n_kpt;
  const int t[NLAYER][2] = {{128, enc + 64}, {128, 128}, {120, 136}, {64, 120}, {64, 128}, {64, 64}, {2, 64}, {24, 128},
                            {16, 4}, {35, 16}, {64, 105}, {32, 64}, {32, 32}, {33, 32}, {32, 32}, {1, 32},
                            {16, 37}, {8, 16}, {1, 8}};
  memcpy(dims, t, sizeof(t));
}

extern "C" int kpn_set_weights(kpn_ctx* c, const kpn_weights* w) {
  if (!c || !w) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  if (w->n_kpt < 1 || w->n_kpt > simt_max_kpt() || w->n_kpt > KPN_MAX_KPT)
    KPN_FAIL(c, KPN_ERR_UNSUPPORTED, "n_kpt=%d unsupported (1..%d)", w->n_kpt, simt_max_kpt());
  if (w->sp_level < 0 || w->sp_level > MAX_SPL || (1 + 2 * w->sp_level) * w->n_kpt + 64 > 232)
    KPN_FAIL(c, KPN_ERR_UNSUPPORTED, "sp_level=%d unsupported", w->sp_level);
  int dims[NLAYER][2];
  expected_dims(w->n_kpt, w->sp_level, dims);
  size_t total = 0;
  DevWeightsF32 hw;
  std::vector<size_t> woff(NLAYER), boff(NLAYER);
  for (int l = 0; l < NLAYER; ++l) {
    const kpn_layer& L = w->layer[l];
    if (!L.w || !L.bias) KPN_FAIL(c, KPN_ERR_ARG, "layer %d: null weight/bias", l);
    if (L.n_out != dims[l][0] || L.n_in != dims[l][1])
      KPN_FAIL(c, KPN_ERR_ARG, "layer %d: shape (%d,%d) != expected (%d,%d)", l, L.n_out, L.n_in, dims[l][0], dims[l][1]);
    hw.K[l] = L.n_in; hw.N[l] = L.n_out; hw.ldw[l] = (L.n_out + 31) / 32 * 32;
    woff[l] = total; total += (size_t)hw.K[l] * hw.ldw[l];
    boff[l] = total; total += (size_t)hw.ldw[l];
  }
  std::vector<float> host(total, 0.0f);
  std::vector<std::vector<float>> We(NLAYER);  // effective weights, row-major [n_out][n_in]
  for (int l = 0; l < NLAYER; ++l) {
    const kpn_layer& L = w->layer[l];
    We[l].resize((size_t)L.n_out * L.n_in);
    for (int o = 0; o < L.n_out; ++o) {
      double scale = 1.0;
      if (L.g) {  // weight norm, dim=0: w = g * v / ||v||_row  (reference src/utils.py:542-543)
        double nn = 0.0;
        for (int i = 0; i < L.n_in; ++i) nn += (double)L.w[(size_t)o * L.n_in + i] * (double)L.w[(size_t)o * L.n_in + i];
        scale = (double)L.g[o] / std::sqrt(nn);
      }
      for (int i = 0; i < L.n_in; ++i) {
        float e = (float)(scale * (double)L.w[(size_t)o * L.n_in + i]);
        We[l][(size_t)o * L.n_in + i] = e;
        host[woff[l] + (size_t)i * hw.ldw[l] + o] = e;
      }
      host[boff[l] + o] = L.bias[o];
    }
  }
  KPN_CUDA(c, cudaDeviceSynchronize());  // the old buffer may still be in use
  KPN_CUDA(c, c->wbuf.reserve(total * sizeof(float)));
  KPN_CUDA(c, cudaMemcpy(c->wbuf.p, host.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  for (int l = 0; l < NLAYER; ++l) {
    hw.wt[l] = c->wbuf.as<float>() + woff[l];
    hw.bias[l] = c->wbuf.as<float>() + boff[l];
  }
  hw.ani_al_abs = std::fabs(w->ani_al);
  KPN_CUDA(c, cudaMemcpy(c->d_wf32, &hw, sizeof(hw), cudaMemcpyHostToDevice));
  // ---- tensor-core engine: fp16 weight tiles in the interleaved core-matrix layout + fp32 constants
  c->tc_weights = false;
  if (tc_supported(3, w->n_kpt, w->sp_level)) {
    const TcPlan plan = make_tc_plan(w->n_kpt);
    std::vector<__half> blob(plan.total_bytes / 2, __float2half_rn(0.0f));
    const size_t geo_bytes = tc_weight_lo_bytes(w->n_kpt);          // bytes of the full W_hi tiles of stages 0..5
    std::vector<__half> pair(tc_pair_blob_bytes(w->n_kpt) / 2, __float2half_rn(0.0f));   // [rank][hi halves | lo halves]
    const bool vseq = vs_run_cols(w->n_kpt) > 0;
    std::vector<__half> pair_vs(vseq ? pair.size() : 0, __float2half_rn(0.0f));           // view-sequential kernel: same tiles, other K permutation
    // stage -> (layer, first row in the tile); stage 4 stacks the density layer 0 and the colour compress layer
    const int stage_layer[TC_NSTAGE] = {L_GEO0, L_GEO1, L_GEO2, L_GEO3, L_DEN0, L_DEN1, L_BASE0, L_BASE1, L_VIS1A, L_VIS1B, L_VIS2A, L_OUT0, L_RE1};
    auto put_pair = [&](std::vector<__half>& dst, int stage, int n, int kk, float wv) {
      // geometry stages: row n of the tile goes to CTA rank n / (Np/2), as row n % (Np/2) of its half tile
      const int Np = plan.st[stage].Np, Nh = Np / 2, rk = n / Nh;
      const size_t half_at = ((size_t)rk * geo_bytes + plan.st[stage].off / 2 + tc::core_offset_bytes(n % Nh, kk, Nh)) / 2;
      const __half hi = __float2half_rn(wv);
      dst[half_at] = hi;
      dst[half_at + geo_bytes / 4] = __float2half_rn(wv - __half2float(hi));   // lo halves follow the hi halves
    };
    auto put_one = [&](int stage, int n, int kk, float wv) {
      const int Np = plan.st[stage].Np;
      const size_t at = (plan.st[stage].off + tc::core_offset_bytes(n, kk, Np)) / 2;
      blob[at] = __float2half_rn(wv);
      if (stage < 6) put_pair(pair, stage, n, kk, wv);
    };
    auto put = [&](int stage, int layer, int row0) {
      const kpn_layer& L = w->layer[layer];
      for (int o = 0; o < L.n_out; ++o) {
        for (int i = 0; i < L.n_in; ++i) {
          put_one(stage, row0 + o, stage < 6 ? tc_kmap(stage, w->n_kpt, i) : i, We[layer][(size_t)o * L.n_in + i]);
          if (vseq && stage < 6) put_pair(pair_vs, stage, row0 + o, tc_kmap_vseq(stage, w->n_kpt, i), We[layer][(size_t)o * L.n_in + i]);
        }
        if (stage < 6 || stage == 12) put_one(stage, row0 + o, tc_kbias(stage, w->n_kpt), L.bias[o]);   // bias row (activation column == 1)
        if (vseq && stage < 6) put_pair(pair_vs, stage, row0 + o, tc_kbias_vseq(stage, w->n_kpt), L.bias[o]);
      }
    };
    for (int sidx = 0; sidx < TC_NSTAGE; ++sidx) put(sidx, stage_layer[sidx], 0);
    put(4, L_CMP, 64);
    KPN_CUDA(c, c->wblob.reserve(plan.total_bytes));
    KPN_CUDA(c, cudaMemcpy(c->wblob.p, blob.data(), plan.total_bytes, cudaMemcpyHostToDevice));
    KPN_CUDA(c, c->wlo.reserve(pair.size() * 2));
    KPN_CUDA(c, cudaMemcpy(c->wlo.p, pair.data(), pair.size() * 2, cudaMemcpyHostToDevice));
    if (vseq) {
      KPN_CUDA(c, c->wlo_vs.reserve(pair_vs.size() * 2));
      KPN_CUDA(c, cudaMemcpy(c->wlo_vs.p, pair_vs.data(), pair_vs.size() * 2, cudaMemcpyHostToDevice));
    }
    TcConsts& T = c->tcc;
    memset(&T, 0, sizeof(T));
    auto bias = [&](int layer, float* dst) { for (int o = 0; o < w->layer[layer].n_out; ++o) dst[o] = w->layer[layer].bias[o]; };
    bias(L_BASE0, T.b_base0); bias(L_BASE1, T.b_base1); bias(L_VIS1A, T.b_vis1a); bias(L_VIS1B, T.b_vis1b);
    bias(L_VIS2A, T.b_vis2a); bias(L_OUT0, T.b_out0);
    for (int o = 0; o < 2; ++o) { for (int i = 0; i < 64; ++i) T.w_p2[o][i] = We[L_DEN2][o * 64 + i]; T.b_p2[o] = w->layer[L_DEN2].bias[o]; }
    for (int o = 0; o < 16; ++o) { for (int i = 0; i < 4; ++i) T.w_re0[o][i] = We[L_RE0][o * 4 + i]; T.b_re0[o] = w->layer[L_RE0].bias[o]; }
    for (int i = 0; i < 32; ++i) T.w_vis2b[i] = We[L_VIS2B][i];
    T.b_vis2b = w->layer[L_VIS2B].bias[0];
    for (int o = 0; o < 8; ++o) { for (int i = 0; i < 16; ++i) T.w_out1[o][i] = We[L_OUT1][o * 16 + i]; T.b_out1[o] = w->layer[L_OUT1].bias[o]; }
    for (int i = 0; i < 8; ++i) T.w_out2[i] = We[L_OUT2][i];
    T.b_out2 = w->layer[L_OUT2].bias[0];
    T.ani_abs = std::fabs(w->ani_al);
    c->tc_weights = true;
  }
  c->n_kpt = w->n_kpt; c->sp_level = w->sp_level; c->sp_scale = w->sp_scale; c->sp_sigma = w->sp_sigma;
  c->have_weights = true;
  return KPN_OK;
}

// ---------------------------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------------------------
static cudaMemcpyKind in_kind(int mem) { return mem == KPN_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice; }

extern "C" int kpn_set_scene(kpn_ctx* c, const kpn_scene* s, void* stream) {
  if (!c || !s) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!c->have_weights) KPN_FAIL(c, KPN_ERR_STATE, "kpn_set_weights must precede kpn_set_scene");
  { int hrc = wd_check(c, st); if (hrc != KPN_OK) return hrc; }
  if (s->n_views < 1 || s->n_views > 3) KPN_FAIL(c, KPN_ERR_UNSUPPORTED, "n_views=%d unsupported (1..3)", s->n_views);
  if (s->n_kpt != c->n_kpt) KPN_FAIL(c, KPN_ERR_ARG, "scene n_kpt=%d != weights n_kpt=%d", s->n_kpt, c->n_kpt);
  if (!s->KRT || !s->extrin || !s->kpt3d || !s->bounds || !s->feat64 || !s->feat8 || !s->feat_tex || !s->img)
    KPN_FAIL(c, KPN_ERR_ARG, "null scene pointer");
  if (s->f64_c != 64 || s->f8_c != 8 || s->ftex_c != 8)
    KPN_FAIL(c, KPN_ERR_UNSUPPORTED, "feature channels (%d,%d,%d) != (64,8,8)", s->f64_c, s->f8_c, s->ftex_c);
  if (s->mem != KPN_MEM_HOST && s->mem != KPN_MEM_DEVICE) KPN_FAIL(c, KPN_ERR_ARG, "bad mem kind");
  const int V = s->n_views;
  cudaMemcpyKind kind = in_kind(s->mem);

  DevScene h;
  memset(&h, 0, sizeof(h));
  h.V = V; h.K = s->n_kpt;
  h.wm1 = s->src_width - 1.0f; h.hm1 = s->src_height - 1.0f;
  h.znear = s->znear; h.zfar = s->zfar;
  h.sdf_invalid = 0.1f / s->nml_scale;
  h.use_fg = s->fg != nullptr;
  h.sp_scale = c->sp_scale;
  h.inv2sig2 = 1.0f / (2.0f * c->sp_sigma * c->sp_sigma);
  h.sp_level = c->sp_level;
  {
    double val = 1.0;
    for (int l = 0; l < c->sp_level; ++l) { h.freq[l] = (float)(3.14159265358979323846 * val); val *= 2.0; }
  }
  struct MapIn { const float* src; int C, H, W, Cp; } maps[4] = {
      {s->feat64, s->f64_c, s->f64_h, s->f64_w, 64}, {s->feat8, s->f8_c, s->f8_h, s->f8_w, 8},
      {s->feat_tex, s->ftex_c, s->ftex_h, s->ftex_w, 8}, {s->img, 3, s->img_h, s->img_w, 4}};
  MapDesc* descs[4] = {&h.f64, &h.f8, &h.ftex, &h.img};
  for (int m = 0; m < 4; ++m) {
    const MapIn& M = maps[m];
    if (M.H < 1 || M.W < 1) KPN_FAIL(c, KPN_ERR_ARG, "map %d has empty extent", m);
    size_t in_bytes = (size_t)V * M.C * M.H * M.W * sizeof(float);
    size_t out_bytes = (size_t)V * M.Cp * M.H * M.W * sizeof(float);
    if (m < 3 && ((s->layout >> m) & 1)) {
      // already [V][H][W][C] (channels-last encoder output, C == Cp): no re-layout pass.  Device maps are gathered from in
      // place (the caller keeps them alive and unchanged until the next kpn_set_scene); host maps are uploaded as they are.
      const void* ptr = M.src;
      if (s->mem == KPN_MEM_HOST) {
        KPN_CUDA(c, c->atlas[m].reserve(out_bytes));
        KPN_CUDA(c, cudaMemcpyAsync(c->atlas[m].p, M.src, in_bytes, cudaMemcpyHostToDevice, st));
        ptr = c->atlas[m].p;
      }
      descs[m]->ptr = ptr; descs[m]->C = M.Cp; descs[m]->H = M.H; descs[m]->W = M.W;
      continue;
    }
    const float* dsrc = M.src;
    if (s->mem == KPN_MEM_HOST) {
      KPN_CUDA(c, c->stage[m].reserve(in_bytes));
      KPN_CUDA(c, cudaMemcpyAsync(c->stage[m].p, M.src, in_bytes, cudaMemcpyHostToDevice, st));
      dsrc = c->stage[m].as<float>();
    }
    KPN_CUDA(c, c->atlas[m].reserve(out_bytes));
    KPN_CUDA(c, launch_pack_nhwc_f32(dsrc, c->atlas[m].as<float>(), V, M.C, M.H, M.W, M.Cp, st));
    c->launches++;
    descs[m]->ptr = c->atlas[m].p; descs[m]->C = M.Cp; descs[m]->H = M.H; descs[m]->W = M.W;
  }
  if (s->fg) {
    size_t bytes = (size_t)V * s->fg_h * s->fg_w;
    KPN_CUDA(c, c->atlas_fg.reserve(bytes));
    KPN_CUDA(c, cudaMemcpyAsync(c->atlas_fg.p, s->fg, bytes, kind, st));
    h.fg.ptr = c->atlas_fg.p; h.fg.C = 1; h.fg.H = s->fg_h; h.fg.W = s->fg_w;
  }
  KPN_CUDA(c, cudaMemcpyAsync(c->d_scene, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  KPN_CUDA(c, cudaMemcpyAsync(c->d_raw_scene->KRT, s->KRT, sizeof(float) * 16 * V, kind, st));
  KPN_CUDA(c, cudaMemcpyAsync(c->d_raw_scene->extrin, s->extrin, sizeof(float) * 16 * V, kind, st));
  KPN_CUDA(c, cudaMemcpyAsync(c->d_raw_scene->kpt3d, s->kpt3d, sizeof(float) * 3 * s->n_kpt, kind, st));
  KPN_CUDA(c, cudaMemcpyAsync(c->d_raw_scene->bounds, s->bounds, sizeof(float) * 6, kind, st));
  KPN_CUDA(c, launch_prep_scene(c->d_raw_scene, c->d_scene, st));
  c->launches++;
  if (s->fg) { KPN_CUDA(c, launch_fg_box(c->d_scene, V, st)); c->launches++; }
  c->have_scene = true;
  c->scene_views = V;
  return KPN_OK;
}

// ---------------------------------------------------------------------------------------------
// shading of a batch of samples with the selected engine
// ---------------------------------------------------------------------------------------------
// Shades the samples of list[0 .. *counter) (device-side count, at most n) with the selected engine; `so` says where the results go.
static int shade_batch(kpn_ctx* c, const SampleSrc& src, const int* list, const int* counter, int* counter2, long long n,
                       int query_mode, const ShadeOut& so, int engine, cudaStream_t st) {
  const bool use_tc = engine != 1;
  if (engine < 0 || engine > 4) KPN_FAIL(c, KPN_ERR_ARG, "kpn_opts.engine = %d (0 .. 4)", engine);
  if (use_tc && !(c->tc_weights && tc_supported(c->scene_views, c->n_kpt, c->sp_level)))
    KPN_FAIL(c, KPN_ERR_UNSUPPORTED, "the tensor-core engine covers n_views == 3, n_kpt in {18, 24}, sp_level == 3; this scene has "
             "n_views=%d n_kpt=%d sp_level=%d: request engine = 1 (fp32 CUDA-core engine, ~25x slower) explicitly",
             c->scene_views, c->n_kpt, c->sp_level);
  cudaEvent_t e0 = nullptr, em = nullptr, e1 = nullptr;   // start | after the geometry kernel | stop
  if (c->profiling) {
    while (c->ev_used + 3 > c->ev_pool.size()) {
      cudaEvent_t a;
      KPN_CUDA(c, cudaEventCreate(&a));
      c->ev_pool.push_back(a);
    }
    e0 = c->ev_pool[c->ev_used]; em = c->ev_pool[c->ev_used + 1]; e1 = c->ev_pool[c->ev_used + 2];
    c->ev_used += 3;
    KPN_CUDA(c, cudaEventRecord(e0, st));
  }
  if (use_tc) {
    KPN_CUDA(c, c->ws_lat.reserve((size_t)n * 48));
    KPN_CUDA(c, c->ws_list2.reserve((size_t)n * 8));
    // geometry kernel: the view-sequential one where it is built (18 keypoints), else the row-per-view one; 3 / 4 force either
    const bool vs_built = vs_run_cols(c->n_kpt) > 0;
    if (engine == 3 && !vs_built)
      KPN_FAIL(c, KPN_ERR_UNSUPPORTED, "engine 3 (view-sequential geometry kernel) is built for n_kpt == 18; this scene has %d", c->n_kpt);
    const bool vs = vs_built && (engine == 0 || engine == 3);
    KPN_CUDA(c, launch_shade_tc(c->d_scene, c->tcc, c->wblob.as<uint8_t>(), vs ? c->wlo_vs.as<uint8_t>() : c->wlo.as<uint8_t>(),
                                engine == 2 ? 0 : 1, vs ? -c->n_kpt : c->n_kpt, src, list, counter, n, query_mode, so, c->ws_lat.p,
                                c->ws_list2.p, counter2, c->num_sms, em, st));
    c->launches += 2;
  }
  else {
    KPN_CUDA(c, launch_shade_simt(c->d_scene, c->d_wf32, src, list, counter, n, query_mode, so, c->num_sms, st));
    if (c->profiling) KPN_CUDA(c, cudaEventRecord(em, st));
    c->launches++;
  }
  if (c->profiling) KPN_CUDA(c, cudaEventRecord(e1, st));
  return KPN_OK;
}

// rays per chunk: bounded by the per-sample workspace; fewer, larger launches amortise the persistent kernels' prologue (weight
// load) and tail (measured 4 / 8 / 32 Mi samples: 30.5 / 30.2 / 30.0 ms per 512x512x128 frame).  KPN_CHUNK_SAMPLES overrides
// the default of 32 Mi samples.
static long long chunk_rays(long long R, int Smax) {
  static const long long chunk_samples = [] {
    const char* e = getenv("KPN_CHUNK_SAMPLES");
    long long v = e ? atoll(e) : 0;
    return v >= (1ll << 16) && v <= (1ll << 28) ? v : (32ll << 20);
  }();
  long long Rc = chunk_samples / Smax;
  Rc = (Rc / 128) * 128;
  if (Rc < 128) Rc = 128;
  return Rc > R ? R : Rc;
}

// every per-chunk workspace buffer of a render of Rc rays x (Sc coarse, Smax total) samples
static int reserve_render(kpn_ctx* c, long long Rc, int Sc, int Smax, bool fine, bool contrib, bool ert) {
  const size_t n = (size_t)Rc * Smax;
  KPN_CUDA(c, c->ws_rayd.reserve((size_t)Rc * 3 * sizeof(float)));
  KPN_CUDA(c, c->ws_raynf.reserve((size_t)Rc * 2 * sizeof(float)));
  KPN_CUDA(c, c->ws_rayseg.reserve((size_t)Rc * 4 * sizeof(int)));   // start | cnt of segment 0, start | cnt of segment 1
  KPN_CUDA(c, c->ws_list.reserve(n * sizeof(int)));
  KPN_CUDA(c, c->ws_ao.reserve(n * sizeof(float2)));
  KPN_CUDA(c, c->ws_rgb.reserve(n * 3 * sizeof(float)));
  KPN_CUDA(c, c->ws_lat.reserve(n * 48));
  KPN_CUDA(c, c->ws_list2.reserve(n * 8));
  if (fine || contrib) KPN_CUDA(c, c->ws_cw.reserve((size_t)Rc * Sc * sizeof(float)));
  if (contrib) KPN_CUDA(c, c->ws_contrib.reserve((size_t)Rc * Sc * sizeof(float)));
  if (fine) KPN_CUDA(c, c->ws_zfine.reserve(n * sizeof(float)));
  if (ert) KPN_CUDA(c, c->ws_ert.reserve((size_t)Rc * sizeof(float)));
  return KPN_OK;
}

extern "C" int kpn_reserve(kpn_ctx* c, long long max_rays, int max_samples) {
  if (!c || max_rays < 1 || max_samples < 1) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  const long long Rc = chunk_rays(max_rays, max_samples);
  int rc = reserve_render(c, Rc, max_samples, max_samples, true, true, true);
  if (rc != KPN_OK) return rc;
  KPN_CUDA(c, c->ws_out.reserve((size_t)max_rays * 11 * sizeof(float)));   // host-output staging: 3+1+1+3+1+1+1 planes
  return KPN_OK;
}

static int begin_counters(kpn_ctx* c, cudaStream_t st) {
  KPN_CUDA(c, cudaMemsetAsync(c->d_counters, 0, sizeof(int) * MAX_CHUNKS, st));
  KPN_CUDA(c, cudaMemsetAsync(c->d_counters2, 0, sizeof(int) * MAX_CHUNKS, st));
  c->counters_used = 0;
  c->last_total = 0;
  return KPN_OK;
}

extern "C" int kpn_render(kpn_ctx* c, const kpn_target* tg, const kpn_opts* op, const kpn_out* out, void* stream) {
  if (!c || !tg || !op || !out) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!c->have_scene) KPN_FAIL(c, KPN_ERR_STATE, "kpn_set_scene must precede kpn_render");
  { int hrc = wd_check(c, st); if (hrc != KPN_OK) return hrc; }
  const int Sc = op->sample_per_ray_c, Sf = op->fine ? op->sample_per_ray_f : 0;
  if (Sc < 1 || Sc > max_coarse_samples()) KPN_FAIL(c, KPN_ERR_ARG, "sample_per_ray_c=%d out of range", Sc);
  if (op->fine && (Sc < 3 || Sf < 1 || Sf > 1024)) KPN_FAIL(c, KPN_ERR_ARG, "fine pass needs S_c>=3 and 1<=S_f<=1024");
  if (tg->nx < 1 || tg->ny < 1 || tg->step < 1 || tg->step_y < 0) KPN_FAIL(c, KPN_ERR_ARG, "empty pixel lattice");
  if (!tg->K || !tg->RT) KPN_FAIL(c, KPN_ERR_ARG, "null target camera");
  const long long R = (long long)tg->nx * tg->ny;
  const int Smax = Sc + Sf;
  if (R * Smax >= (1ll << 40)) KPN_FAIL(c, KPN_ERR_ARG, "too many samples");

  // target camera -> device constants
  cudaMemcpyKind kind = in_kind(tg->mem);
  KPN_CUDA(c, cudaMemcpyAsync(c->d_raw_target->K, tg->K, sizeof(float) * 16, kind, st));
  KPN_CUDA(c, cudaMemcpyAsync(c->d_raw_target->RT, tg->RT, sizeof(float) * 16, kind, st));
  DevTarget ht;
  memset(&ht, 0, sizeof(ht));
  ht.znear = tg->znear; ht.zfar = tg->zfar; ht.x0 = tg->x0; ht.y0 = tg->y0; ht.step = tg->step; ht.nx = tg->nx; ht.ny = tg->ny;
  ht.step_y = tg->step_y > 0 ? tg->step_y : tg->step;
  KPN_CUDA(c, cudaMemcpyAsync(c->d_target, &ht, sizeof(ht), cudaMemcpyHostToDevice, st));
  KPN_CUDA(c, launch_prep_target(c->d_raw_target, c->d_target, c->d_scene, st));
  c->launches++;
  const float* d_o = reinterpret_cast<const float*>(reinterpret_cast<const char*>(c->d_target) + offsetof(DevTarget, o));

  // outputs: render into device planes, copy to the host at the end if requested
  float* user[7] = {out->tex_fg, out->depth, out->alpha, out->tex_fg_fine, out->depth_fine, out->alpha_fine, out->sdf};
  const int planes[7] = {3, 1, 1, 3, 1, 1, 1};
  float* dev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const bool host_out = out->mem == KPN_MEM_HOST;
  if (host_out) {
    size_t tot = 0;
    for (int i = 0; i < 7; ++i) if (user[i]) tot += (size_t)planes[i] * R;
    KPN_CUDA(c, c->ws_out.reserve(tot * sizeof(float)));
    size_t off = 0;
    for (int i = 0; i < 7; ++i) if (user[i]) { dev[i] = c->ws_out.as<float>() + off; off += (size_t)planes[i] * R; }
  } else {
    for (int i = 0; i < 7; ++i) dev[i] = user[i];
  }

  const long long Rc = chunk_rays(R, Smax);
  const long long nchunks = (R + Rc - 1) / Rc;
  if (nchunks * 4 > MAX_CHUNKS) KPN_FAIL(c, KPN_ERR_ARG, "frame too large for one call (%lld chunks)", nchunks);
  const bool need_contrib = op->fine || out->contrib;
  const bool ert_on = op->ert_eps > 0.0f;
  int rc = reserve_render(c, Rc, Sc, Smax, op->fine != 0, need_contrib, ert_on);
  if (rc != KPN_OK) return rc;
  rc = begin_counters(c, st);
  if (rc != KPN_OK) return rc;
  cudaMemcpyKind okind = host_out ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  int* seg = c->ws_rayseg.as<int>();
  int* list = c->ws_list.as<int>();
  float2* ao = c->ws_ao.as<float2>();
  float* rgbw = c->ws_rgb.as<float>();
  // One pass over nr rays x S samples (zbuf == nullptr: the coarse pass, uniform depths): front (rays + validity + ray-ordered
  // work list) -> shading -> compositing.  Early-ray termination (ert_eps > 0, S >= 8): the front half of every ray is shaded
  // and composited first; rays whose transmittance behind it is < ert_eps get no entries for the back half (their remaining
  // contribution to any channel is < ert_eps; the reference has no such option, ert_eps = 0 reproduces it exactly).  The
  // back half's entries live in the list range [nr*half, nr*S) and have their own per-ray ranges (segment 1).
  auto pass = [&](int r0, int nr, int S, const float* zbuf, int slot, float* color, float* depth, float* alpha, float* sdf,
                  float* cw) -> int {
    SampleSrc src;
    memset(&src, 0, sizeof(src));
    src.mode = 0; src.S = S; src.ray_d = c->ws_rayd.as<float>(); src.z = zbuf; src.ray_nf = c->ws_raynf.as<float>(); src.o = d_o;
    const long long n = (long long)nr * S;
    int* start0 = seg; int* cnt0 = seg + Rc; int* start1 = seg + 2 * Rc; int* cnt1 = seg + 3 * Rc;
    const bool two = ert_on && S >= 8;
    const int half = S / 2;
    ShadeOut so;
    memset(&so, 0, sizeof(so));
    so.ao = ao; so.rgb = rgbw;
    KPN_CUDA(c, launch_front(c->d_scene, c->d_target, r0, nr, S, zbuf, c->ws_rayd.as<float>(), c->ws_raynf.as<float>(), list, 0,
                             c->d_counters + slot, start0, cnt0, two ? ErtSegment{0, half, nullptr, 0.0f} : ErtSegment{0, 0, nullptr, 0.0f}, st));
    c->launches++;
    so.list_base = 0;
    int r1 = shade_batch(c, src, list, c->d_counters + slot, c->d_counters2 + slot, two ? (long long)nr * half : n, 0, so, op->engine, st);
    if (r1 != KPN_OK) return r1;
    if (two) {
      KPN_CUDA(c, launch_composite(list, ao, rgbw, start0, cnt0, nullptr, nullptr, zbuf, c->ws_raynf.as<float>(), r0, nr, S, R, nullptr,
                                   nullptr, nullptr, nullptr, c->ws_ert.as<float>(), nullptr, st));
      const int base1 = nr * half;
      KPN_CUDA(c, launch_front(c->d_scene, c->d_target, r0, nr, S, zbuf ? zbuf : nullptr, c->ws_rayd.as<float>(), c->ws_raynf.as<float>(),
                               list, base1, c->d_counters + slot + 1, start1, cnt1, ErtSegment{half, S, c->ws_ert.as<float>(), op->ert_eps}, st));
      c->launches += 2;
      so.list_base = base1;
      r1 = shade_batch(c, src, list + base1, c->d_counters + slot + 1, c->d_counters2 + slot + 1, n - base1, 0, so, op->engine, st);
      if (r1 != KPN_OK) return r1;
    }
    KPN_CUDA(c, launch_composite(list, ao, rgbw, start0, cnt0, two ? start1 : nullptr, two ? cnt1 : nullptr, zbuf, c->ws_raynf.as<float>(),
                                 r0, nr, S, R, color, depth, alpha, sdf, nullptr, cw, st));
    c->launches++;
    return KPN_OK;
  };

  for (long long ch = 0; ch < nchunks; ++ch) {
    const long long r0 = ch * Rc;
    const int nr = (int)((R - r0) < Rc ? (R - r0) : Rc);
    // coarse pass; with early-ray termination its weights must still cover every sample of the ray for the resampling pass:
    // rays that terminated early simply have zero weight behind the cut
    rc = pass((int)r0, nr, Sc, nullptr, (int)(4 * ch), dev[0], dev[1], dev[2], nullptr, need_contrib ? c->ws_cw.as<float>() : nullptr);
    if (rc != KPN_OK) return rc;
    c->last_total += (unsigned long long)nr * Sc;
    if (need_contrib) {
      // dense weights (debug output) and / or the resampled + merged depths of the fine pass, one warp per ray
      const bool two = ert_on && Sc >= 8;
      float* zout = (op->fine && !op->z_fine_override) ? c->ws_zfine.as<float>() : nullptr;
      float* cden = out->contrib ? c->ws_contrib.as<float>() : nullptr;
      if (zout || cden) {
        KPN_CUDA(c, launch_resample(list, c->ws_cw.as<float>(), seg, seg + Rc, two ? seg + 2 * Rc : nullptr, two ? seg + 3 * Rc : nullptr,
                                    c->ws_raynf.as<float>(), nr, Sc, zout ? Sf : 0, zout, cden, st));
        c->launches++;
      }
      if (out->contrib)
        KPN_CUDA(c, cudaMemcpyAsync(out->contrib + r0 * Sc, c->ws_contrib.p, (size_t)nr * Sc * sizeof(float), okind, st));
    }
    if (op->fine) {
      if (op->z_fine_override)
        KPN_CUDA(c, cudaMemcpyAsync(c->ws_zfine.p, op->z_fine_override + r0 * Smax, (size_t)nr * Smax * sizeof(float),
                                    in_kind(out->mem), st));
      rc = pass((int)r0, nr, Smax, c->ws_zfine.as<float>(), (int)(4 * ch + 2), dev[3], dev[4], dev[5], dev[6], nullptr);
      if (rc != KPN_OK) return rc;
      c->last_total += (unsigned long long)nr * Smax;
      if (out->z_fine)
        KPN_CUDA(c, cudaMemcpyAsync(out->z_fine + r0 * Smax, c->ws_zfine.p, (size_t)nr * Smax * sizeof(float), okind, st));
    }
  }
  c->counters_used = (int)(4 * nchunks);
  if (host_out) {
    for (int i = 0; i < 7; ++i)
      if (user[i]) KPN_CUDA(c, cudaMemcpyAsync(user[i], dev[i], (size_t)planes[i] * R * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  return wd_snapshot(c, st);
}

extern "C" int kpn_query(kpn_ctx* c, const float* pts, const float* view, int n, float* out5, uint8_t* valid, int mem,
                         const kpn_opts* op, void* stream) {
  if (!c || !pts || !view || !out5 || n < 0) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!c->have_scene) KPN_FAIL(c, KPN_ERR_STATE, "kpn_set_scene must precede kpn_query");
  { int hrc = wd_check(c, st); if (hrc != KPN_OK) return hrc; }
  int rc = begin_counters(c, st);
  if (rc != KPN_OK) return rc;
  if (n == 0) return KPN_OK;
  const float* dp = pts; const float* dv = view;
  float* dout = out5; uint8_t* dvalid = valid;
  if (mem == KPN_MEM_HOST) {
    size_t fb = (size_t)n * 3 * sizeof(float);
    KPN_CUDA(c, c->ws_in.reserve(2 * fb));
    KPN_CUDA(c, cudaMemcpyAsync(c->ws_in.p, pts, fb, cudaMemcpyHostToDevice, st));
    KPN_CUDA(c, cudaMemcpyAsync(c->ws_in.as<char>() + fb, view, fb, cudaMemcpyHostToDevice, st));
    dp = c->ws_in.as<float>(); dv = reinterpret_cast<const float*>(c->ws_in.as<char>() + fb);
    KPN_CUDA(c, c->ws_out.reserve((size_t)n * 5 * sizeof(float) + (size_t)n));
    dout = c->ws_out.as<float>();
    dvalid = reinterpret_cast<uint8_t*>(c->ws_out.as<char>() + (size_t)n * 5 * sizeof(float));
  } else if (!dvalid) {
    KPN_CUDA(c, c->ws_out.reserve((size_t)n));
    dvalid = c->ws_out.as<uint8_t>();
  }
  const long long chunk = 4ll << 20;
  const long long nchunks = (n + chunk - 1) / chunk;
  if (nchunks > MAX_CHUNKS) KPN_FAIL(c, KPN_ERR_ARG, "too many points");
  for (long long ch = 0; ch < nchunks; ++ch) {
    long long off = ch * chunk;
    long long m = (n - off) < chunk ? (n - off) : chunk;
    SampleSrc src;
    memset(&src, 0, sizeof(src));
    src.mode = 1; src.pts = dp + 3 * off; src.view = dv + 3 * off;
    KPN_CUDA(c, c->ws_list.reserve((size_t)m * sizeof(int)));
    KPN_CUDA(c, launch_compact(c->d_scene, src, m, c->ws_list.as<int>(), c->d_counters + ch, dout + 5 * off, dvalid + off, st));
    c->launches++;
    ShadeOut so;
    memset(&so, 0, sizeof(so));
    so.out5 = dout + 5 * off;
    rc = shade_batch(c, src, c->ws_list.as<int>(), c->d_counters + ch, c->d_counters2 + ch, m, 1, so, op ? op->engine : 0, st);
    if (rc != KPN_OK) return rc;
  }
  c->counters_used = (int)nchunks;
  c->last_total = (unsigned long long)n;
  if (mem == KPN_MEM_HOST) {
    KPN_CUDA(c, cudaMemcpyAsync(out5, dout, (size_t)n * 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (valid) KPN_CUDA(c, cudaMemcpyAsync(valid, dvalid, (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  return wd_snapshot(c, st);
}

extern "C" int kpn_get_stats(kpn_ctx* c, kpn_stats* stats, void* stream) {
  if (!c || !stats) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int> h(c->counters_used > 0 ? c->counters_used : 1, 0), h2(c->counters_used > 0 ? c->counters_used : 1, 0);
  if (c->counters_used > 0) {
    KPN_CUDA(c, cudaMemcpyAsync(h.data(), c->d_counters, sizeof(int) * c->counters_used, cudaMemcpyDeviceToHost, st));
    KPN_CUDA(c, cudaMemcpyAsync(h2.data(), c->d_counters2, sizeof(int) * c->counters_used, cudaMemcpyDeviceToHost, st));
  }
  KPN_CUDA(c, tc_watchdog_read_async(c->h_wd, st));
  KPN_CUDA(c, cudaStreamSynchronize(st));
  unsigned long long valid = 0, coloured = 0;
  for (int i = 0; i < c->counters_used; ++i) { valid += (unsigned long long)h[i]; coloured += (unsigned long long)h2[i]; }
  stats->samples_total = c->last_total;
  stats->samples_valid = valid;
  stats->samples_coloured = coloured;
  stats->kernel_launches = c->launches;
  double ms = 0.0, gms = 0.0;
  for (size_t i = 0; i + 2 < c->ev_used; i += 3) {   // triples: start, after the geometry kernel, stop
    float t = 0.0f, tg = 0.0f;
    KPN_CUDA(c, cudaEventElapsedTime(&t, c->ev_pool[i], c->ev_pool[i + 2]));
    KPN_CUDA(c, cudaEventElapsedTime(&tg, c->ev_pool[i], c->ev_pool[i + 1]));
    ms += (double)t;
    gms += (double)tg;
  }
  stats->shade_launches = c->ev_used / 3;
  stats->shade_ms = ms;
  stats->geo_ms = gms;
  c->ev_used = 0;
  return wd_check(c, st);
}

// Debug hook.  out16[0..7] receive the device watchdog words of the tensor-core kernels (out16[0] != 0: a barrier wait gave
// up at block out16[1], thread out16[2], tag out16[3]); enable != 0 clears them afterwards.
extern "C" int kpn_debug_timing(kpn_ctx* c, int enable, unsigned long long* out16) {
  if (!c) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  KPN_CUDA(c, cudaDeviceSynchronize());
  unsigned int wd[8];
  KPN_CUDA(c, tc_watchdog_read(wd, enable != 0));
  if (out16) { for (int i = 0; i < 16; ++i) out16[i] = i < 8 ? wd[i] : 0ull; }
  return KPN_OK;
}

// Host-only test hook: the K index (element of the fp16 activation row) each input of a geometry stage is multiplied with, and
// the K index of its bias row (tests/test_host_abi.py checks that the layer-0 permutation is a bijection).
extern "C" int kpn_debug_kmap(int stage, int n_kpt, int n_inputs, int* kmap_out, int* kbias_out, int* kpad_out) {
  if (stage < 0 || stage > 5 || !tc_supported(3, n_kpt, 3) || n_inputs < 0 || !kmap_out) return KPN_ERR_ARG;
  for (int i = 0; i < n_inputs; ++i) kmap_out[i] = tc_kmap(stage, n_kpt, i);
  if (kbias_out) *kbias_out = tc_kbias(stage, n_kpt);
  if (kpad_out) *kpad_out = make_tc_plan(n_kpt).st[stage].Kp;
  return KPN_OK;
}

// Debug, instrumented build (-DKPN_STAGE_TIMING) only: cycle stamps of one issuer warp of the geometry kernel (layout in
// kpn_shade_tc.cu); out receives up to n_words values, *n_tiles the tiles recorded since the last call.
extern "C" int kpn_debug_stage_times(kpn_ctx* c, unsigned long long* out, int n_words, int* n_tiles) {
  if (!c || !out || !n_tiles) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  cudaError_t e = tc_stage_times(out, n_words, n_tiles);
  if (e == cudaErrorNotSupported) KPN_FAIL(c, KPN_ERR_UNSUPPORTED, "library was not built with -DKPN_STAGE_TIMING");
  KPN_CUDA(c, e);
  return KPN_OK;
}

extern "C" int kpn_decode_views(kpn_ctx* c, const uint8_t* images, const uint8_t* masks, const double* cams, int n_views, int src_h,
                                int src_w, int factor, float* out_img, uint8_t* out_mask, int mem, void* stream) {
  if (!c || !images || !cams || !out_img) return KPN_ERR_ARG;
  DeviceGuard g(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_views < 1 || src_h < 1 || src_w < 1 || factor < 1 || src_h % factor || src_w % factor)
    KPN_FAIL(c, KPN_ERR_ARG, "decode: %d views of %dx%d by 1/%d", n_views, src_h, src_w, factor);
  if (mem != KPN_MEM_HOST && mem != KPN_MEM_DEVICE) KPN_FAIL(c, KPN_ERR_ARG, "bad mem kind");
  const size_t npx = (size_t)n_views * src_h * src_w, nout = npx / ((size_t)factor * factor);
  const size_t cam_bytes = (size_t)n_views * 18 * sizeof(double);
  const uint8_t* d_img = images; const uint8_t* d_msk = masks; const void* d_cam = cams;
  float* d_out = out_img; uint8_t* d_om = out_mask;
  if (mem == KPN_MEM_HOST) {
    // staging layout: cams | images | masks | out_img | out_mask
    const size_t off_img = (cam_bytes + 255) / 256 * 256, off_msk = off_img + (npx * 3 + 255) / 256 * 256;
    const size_t off_out = off_msk + (npx + 255) / 256 * 256, off_om = off_out + nout * 3 * sizeof(float);
    KPN_CUDA(c, c->ws_in.reserve(off_om + nout));
    char* b = c->ws_in.as<char>();
    KPN_CUDA(c, cudaMemcpyAsync(b, cams, cam_bytes, cudaMemcpyHostToDevice, st));
    KPN_CUDA(c, cudaMemcpyAsync(b + off_img, images, npx * 3, cudaMemcpyHostToDevice, st));
    if (masks) KPN_CUDA(c, cudaMemcpyAsync(b + off_msk, masks, npx, cudaMemcpyHostToDevice, st));
    d_cam = b; d_img = reinterpret_cast<uint8_t*>(b + off_img); d_msk = masks ? reinterpret_cast<uint8_t*>(b + off_msk) : nullptr;
    d_out = reinterpret_cast<float*>(b + off_out); d_om = out_mask ? reinterpret_cast<uint8_t*>(b + off_om) : nullptr;
  }
  KPN_CUDA(c, launch_decode_views(d_img, d_msk, d_cam, n_views, src_h, src_w, factor, d_out, d_om, st));
  c->launches++;
  if (mem == KPN_MEM_HOST) {
    KPN_CUDA(c, cudaMemcpyAsync(out_img, d_out, nout * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out_mask) KPN_CUDA(c, cudaMemcpyAsync(out_mask, d_om, nout, cudaMemcpyDeviceToHost, st));
  }
  return KPN_OK;
}

extern "C" int kpn_set_profiling(kpn_ctx* c, int enable) {
  if (!c) return KPN_ERR_ARG;
  c->profiling = enable != 0;
  return KPN_OK;
}
