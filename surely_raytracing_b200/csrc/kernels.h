// kernels.h -- launch wrappers implemented in kernels.cu, called by api.cpp
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtb200.h"
#include "device_scene.h"

namespace rtb {

// accumulation buffer: 4 x unsigned long long per pixel, see device_scene.h ACCUM_SCALE

// one persistent launch: per-lane path regeneration over [s_begin, s_end) for every pixel
cudaError_t launch_render_mega(const DScene& S, int64_t s_begin, int64_t s_end, unsigned long long* d_accum, DStats* d_stats,
                               bool collect_stats, cudaStream_t stream, int* launches);

// wavefront pipeline: ray-gen / extend / shade+accumulate kernels over SoA ray queues
struct WavefrontContext {   // per-scene host state of the wavefront driver (filled by api.cpp)
  void* host_counters;      // pinned mirror of the device counters (>= wavefront_counters_bytes())
  int sms;
};
size_t wavefront_counters_bytes();
struct WavefrontOptions {   // rtb_scene_set_option (include/rtb200.h RTB_OPT_*); zero-initialised = defaults, except where noted
  int64_t capacity;         // path slots of the queues (0: sized by the call)
  int exact_leaves;         // 1: the f64 primitive tests run inside the traversal (the round-1 kernel; A/B arm)
  int smem_top;             // 1: the first BVH levels are staged in shared memory (A/B arm)
  int no_defer_rare;        // 1: textured Lambertian items are shaded in place instead of by k_wf_shade_rare
  int extend_blocks_per_sm; // cap of the persistent extend grid (0: what fits)
  int finish_below;         // rays left at which k_wf_finish takes over (-1: default 65536, 0: never)
  int profile;              // 1: per-stage CUDA-event totals on stderr, 2: per iteration
  int defer_rare;           // derived: !no_defer_rare
};
size_t wavefront_workspace_bytes(const DScene& S, int64_t paths_per_wave);
cudaError_t launch_render_wavefront(const DScene& S, const WavefrontContext& ctx, const WavefrontOptions& opt, int64_t s_begin, int64_t s_end,
                                    unsigned long long* d_accum, DStats* d_stats, bool collect_stats, void* d_workspace,
                                    size_t workspace_bytes, int64_t paths_per_wave, cudaStream_t stream, int* launches,
                                    double* stage_ms /* [3], filled when opt.profile */);
cudaError_t launch_add_count(unsigned long long* d_accum, int64_t n_pixels, unsigned long long n_strata, cudaStream_t stream);
// rtb_trace through the pipeline's own kernels (queue records, k_wf_extend, k_wf_extend_exact, exact resolution)
cudaError_t launch_trace_wavefront(const DScene& S, const WavefrontContext& ctx, const WavefrontOptions& opt, const RtbRay* d_rays, int64_t n,
                                   bool secondary, RtbHit* d_hits, int* d_cand_counts, void* d_workspace, size_t workspace_bytes,
                                   unsigned long long* overflows, cudaStream_t stream);

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
// fixed-point sums -> f64 radiance sums (poisoned pixels: NaN); add = 1 accumulates INTO d_pixels_rgb (Q24)
cudaError_t launch_accum_to_f64(const unsigned long long* d_accum, int64_t n_pixels, double* d_pixels_rgb, int add, cudaStream_t stream);
cudaError_t launch_eval_dielectric(const double* d_in, int64_t n, double* d_out, cudaStream_t stream);
// Random123 known-answer hook: out[4*i..] = philox4x32_10(ctr[4*i..], key[2*i..]) evaluated by the device code
cudaError_t launch_philox(const uint32_t* d_ctr_key, int64_t n, uint32_t* d_out, cudaStream_t stream);

}  // namespace rtb
