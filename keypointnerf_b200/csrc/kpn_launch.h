// Host-side launcher declarations (implemented next to their kernels).
#pragma once
#include "kpn_types.cuh"
#include "kpn_tc_types.cuh"

namespace kpn {

cudaError_t launch_pack_nhwc_f32(const float* in, float* out, int V, int C, int H, int W, int Cp, cudaStream_t st);
cudaError_t launch_prep_scene(const RawScene* raw, DevScene* sc, cudaStream_t st);
cudaError_t launch_fg_box(DevScene* sc, int n_views, cudaStream_t st);
cudaError_t launch_prep_target(const RawTarget* raw, DevTarget* tg, const DevScene* sc, cudaStream_t st);
cudaError_t launch_front(const DevScene* sc, const DevTarget* tg, int r0, int nr, int S, const float* zbuf, float* ray_d,
                         float* ray_nf, int* list, int list_base, int* counter, int* ray_start, int* ray_cnt, const ErtSegment& ert,
                         cudaStream_t st);
cudaError_t launch_compact(const DevScene* sc, const SampleSrc& src, long long n, int* list, int* counter, float* out5,
                           uint8_t* valid_out, cudaStream_t st);
cudaError_t launch_shade_simt(const DevScene* sc, const DevWeightsF32* W, const SampleSrc& src, const int* list,
                              const int* counter, long long n_max, int query_mode, const ShadeOut& so, int num_sms,
                              cudaStream_t st);
cudaError_t launch_composite(const int* list, const float2* ao, const float* rgb, const int* start0, const int* cnt0,
                             const int* start1, const int* cnt1, const float* zbuf, const float* ray_nf, int r0, int nr, int S,
                             long long plane, float* color, float* depth, float* alpha, float* sdf, float* ray_alpha, float* cw,
                             cudaStream_t st);
cudaError_t launch_resample(const int* list, const float* cw, const int* start0, const int* cnt0, const int* start1, const int* cnt1,
                            const float* ray_nf, int nr, int Sc, int Sf, float* zout, float* contrib_out, cudaStream_t st);
// Tensor-core engine: geometry+density kernel (CTA pairs), then the colour kernel on the samples with density > 0.
// wblob: full fp16 W_hi tiles (colour stages are read from it); wpair: per-CTA-rank half-blobs [W_hi halves | W_lo halves] of the
// geometry stages; lat_scratch: n_max x 48 bytes, list2: n_max x int2, count2: device int (zeroed by the caller).
cudaError_t launch_shade_tc(const DevScene* sc, const TcConsts& C, const uint8_t* wblob, const uint8_t* wpair, int two_term, int n_kpt,
                            const SampleSrc& src, const int* list, const int* counter, long long n_max, int query_mode,
                            const ShadeOut& so, void* lat_scratch, void* list2, int* count2, int num_sms, cudaEvent_t after_geo,
                            cudaStream_t st);
size_t tc_pair_blob_bytes(int n_kpt);
cudaError_t tc_watchdog_read(unsigned int out[8], bool reset);
cudaError_t tc_watchdog_read_async(unsigned int* pinned_out8, cudaStream_t st);   // stream-ordered copy into pinned host memory
cudaError_t tc_watchdog_clear_async(cudaStream_t st);
cudaError_t tc_stage_times(unsigned long long* out, int n_words, int* n_tiles);   // instrumented build only
size_t tc_weight_blob_bytes(int n_kpt);
size_t tc_weight_lo_bytes(int n_kpt);
bool tc_supported(int n_views, int n_kpt, int sp_level);
// source-view decode (kpn_decode.cu): cams = V x {double ir[9], fx, fy, cx, cy, k1, k2, p1, p2, k3} on the device
cudaError_t launch_decode_views(const uint8_t* images, const uint8_t* masks, const void* cams, int V, int H0, int W0, int factor,
                                float* out_img, uint8_t* out_mask, cudaStream_t st);
int max_coarse_samples();
int simt_max_kpt();

}  // namespace kpn
