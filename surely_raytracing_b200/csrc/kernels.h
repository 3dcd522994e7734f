// kernels.h -- launch wrappers implemented in kernels.cu, called by api.cpp
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtb200.h"
#include "device_scene.h"

namespace rtb {

// one persistent launch: per-lane path regeneration over [s_begin, s_end) for every pixel
cudaError_t launch_render_mega(const DScene& S, int64_t s_begin, int64_t s_end, float4* d_accum, DStats* d_stats,
                               bool collect_stats, cudaStream_t stream, int* launches);

// wavefront pipeline: ray-gen / extend / shade+accumulate kernels over SoA ray queues
constexpr int WF_MAX_SUB = 4;
struct WavefrontContext {   // per-scene host state of the wavefront driver
  void* host_counters;      // pinned mirrors of the device counters (one per sub-pipeline)
  int sms;
  int extend_blocks_per_sm[2];
  int pool_blocks_per_sm[2];
  int extend_kind;          // 0 = per-lane rays (k_wf_extend), 1 = shared-memory ray pool (k_wf_extend_pool)
  int defer_rare;           // 1 = textured Lambertian items are shaded by k_wf_shade_rare (dense), the main kernel has no texture code
  int shade_tma;            // 1 = persistent shade kernel with TMA-staged tiles (k_wf_shade_tma)
  int shade_tma_blocks_per_sm;
  int n_sub;                // sub-pipelines (streams) the stratum range is split over
  cudaStream_t streams[WF_MAX_SUB];
  cudaEvent_t ev_done[WF_MAX_SUB];
  cudaEvent_t ev_start;
};
cudaError_t wavefront_context_create(WavefrontContext* ctx);
void wavefront_context_destroy(WavefrontContext* ctx);
size_t wavefront_workspace_bytes(const DScene& S, int64_t paths_per_wave);
cudaError_t launch_render_wavefront(const DScene& S, const WavefrontContext& ctx, int64_t s_begin, int64_t s_end,
                                    float4* d_accum, DStats* d_stats, bool collect_stats, void* d_workspace,
                                    size_t workspace_bytes, int64_t paths_per_wave, cudaStream_t stream, int* launches);

size_t trace_scratch_bytes(int64_t n);
cudaError_t launch_trace(const DScene& S, const RtbRay* d_rays, int64_t n, uint32_t flags, RtbHit* d_hits,
                         void* d_scratch, cudaStream_t stream);
cudaError_t launch_medium_interval(const DScene& S, int medium, const RtbRay* d_rays, int64_t n, double* d_t0,
                                   double* d_t1, cudaStream_t stream);
cudaError_t launch_eval_texture(const DScene& S, int texture, const double* d_uvp, int64_t n, double* d_rgb,
                                cudaStream_t stream);
cudaError_t launch_eval_light_pdf(const DScene& S, const double* d_od, int64_t n, double* d_pdf, cudaStream_t stream);
cudaError_t launch_write_color(const double* d_pixels, int64_t n_values, double spp, double exposure, uint8_t* d_out,
                               cudaStream_t stream);
cudaError_t launch_accum_to_f64(const float4* d_accum, int64_t n_pixels, double* d_pixels_rgb, cudaStream_t stream);

}  // namespace rtb
