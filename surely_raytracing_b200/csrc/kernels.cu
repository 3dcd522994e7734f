// kernels.cu -- sm_100a kernels of the hot path and their launch wrappers.
//
//   k_render_mega       persistent per-pixel integrator with per-lane path regeneration (A/B arm)
//   k_wf_*              wavefront pipeline: ray-gen -> extend -> shade/accumulate over SoA queues
//   k_trace & friends   deterministic parity harness (closest hit, medium intervals, KAT hooks)
//   k_write_color       output stage (reference src/color.rs:8-33)
//
// No tensor cores anywhere: nothing on this path is a dense contraction.
#include <cuda_runtime.h>

#include "kernels.h"
#include "rtb_device.cuh"

namespace rtb {

// ------------------------------------------------------------------------------------------------
// pixel <-> thread mapping: each warp owns an 8x4 pixel tile (coherent primary rays, similar
// path lengths inside a warp).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tile_pixel(const DCamera& cam, long long warp_global, int lane, uint32_t& pixel) {
  const int tiles_x = (cam.width + 7) >> 3;
  const int tx = (int)(warp_global % tiles_x), ty = (int)(warp_global / tiles_x);
  const int x = tx * 8 + (lane & 7), y = ty * 4 + (lane >> 3);
  pixel = (uint32_t)(y * cam.width + x);
  return x < cam.width && y < cam.height;
}

__device__ __forceinline__ void flush_stats(const DStats& local, DStats* global) {
  atomicAdd(&global->paths, local.paths);
  atomicAdd(&global->exact_tests, local.prim_tests);  // every test of this kernel is an exact one
  atomicAdd(&global->segments, local.segments);
  atomicAdd(&global->node_visits, local.node_visits);
  atomicAdd(&global->prim_tests, local.prim_tests);
  atomicAdd(&global->medium_probes, local.medium_probes);
  atomicAdd(&global->nonfinite, local.nonfinite);
}

// ------------------------------------------------------------------------------------------------
// megakernel: the loop of render_par_lights (reference src/render.rs:179-191) with one thread per
// pixel.  A lane whose path ends immediately starts its next stratum, so every lane of the warp
// keeps tracing segments until its pixel's sample range is exhausted (per-lane regeneration).
// The pixel's sum is kept in registers as the same 64-bit fixed point the accumulation buffer holds and added
// once, by the one thread that owns the pixel: no atomics, and the same bits as any other split of the samples.
// ------------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) k_render_mega(const __grid_constant__ DScene S, long long s_begin,
                                                     long long s_end, unsigned long long* __restrict__ accum,
                                                     DStats* __restrict__ stats) {
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t pixel;
  if (!tile_pixel(S.cam, warp_global, threadIdx.x & 31, pixel)) return;
  DStats st = {};
  unsigned long long sum_r = 0, sum_g = 0, sum_b = 0, poison = 0;
  long long s = s_begin;
  PathState ps;
  bool alive = false;
  float Lr = 0.f, Lg = 0.f, Lb = 0.f;
  RTB_LOOP_ENTER();
  for (;;) {
    if (!alive) {
      if (s >= s_end) break;
      generate_primary(S, pixel, (uint32_t)s, ps);
      s++;
      Lr = Lg = Lb = 0.f;
      alive = true;
      if (STATS) st.paths++;
    }
    Event ev;
    if (STATS) st.segments++;
    extend<STATS>(S, ps, ev, &st);
    alive = shade(S, ps, ev, Lr, Lg, Lb, &st, STATS);
    if (!alive) {
      const bool finite = (fabsf(Lr) < 3.0e38f) && (fabsf(Lg) < 3.0e38f) && (fabsf(Lb) < 3.0e38f);
      if (finite) {
        sum_r += accum_fixed(Lr); sum_g += accum_fixed(Lg); sum_b += accum_fixed(Lb);
      } else if (S.flags & 2u) {
        poison = ACCUM_POISON;
      } else if (STATS) {
        st.nonfinite++;
      }
    }
  }
  RTB_LOOP_LEAVE();
  unsigned long long* a = accum + 4ull * pixel;
  a[0] += sum_r; a[1] += sum_g; a[2] += sum_b;
  a[3] = (a[3] + (unsigned long long)(s_end - s_begin)) | poison;
  if (STATS) flush_stats(st, stats);
}

cudaError_t launch_render_mega(const DScene& S, int64_t s_begin, int64_t s_end, unsigned long long* d_accum, DStats* d_stats,
                               bool collect_stats, cudaStream_t stream, int* launches) {
  const long long tiles = (long long)((S.cam.width + 7) / 8) * ((S.cam.height + 3) / 4);
  const int block = 128;
  const long long blocks = (tiles * 32 + block - 1) / block;
  if (collect_stats) k_render_mega<true><<<(unsigned)blocks, block, 0, stream>>>(S, s_begin, s_end, d_accum, d_stats);
  else k_render_mega<false><<<(unsigned)blocks, block, 0, stream>>>(S, s_begin, s_end, d_accum, d_stats);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// deterministic parity harness
// ------------------------------------------------------------------------------------------------
// Two kernels: (1) closest hit -> compact record {prim, t, alpha, beta}; (2) completion of the full
// f64 hit record (point, normal, uv, face) for the winners.
struct TraceHit { double t, a, b; int prim, pad; };

__device__ __forceinline__ Ray load_ray(const RtbRay& in) {
  Ray r;
  r.ox = in.origin[0]; r.oy = in.origin[1]; r.oz = in.origin[2];
  r.dx = in.direction[0]; r.dy = in.direction[1]; r.dz = in.direction[2];
  r.time = in.time;
  return r;
}

__global__ void __launch_bounds__(128) k_trace_closest(const __grid_constant__ DScene S, const RtbRay* __restrict__ rays,
                                                       long long n, uint32_t flags, TraceHit* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RtbRay in = rays[i];
  const Ray r = load_ray(in);
  Hit best;
  hit_reset(best);
  if (flags & RTB_TRACE_BRUTE_FORCE) closest_surface_brute(S, r, in.t_min, best);
  else if (S.n_surface_prims > 0) closest_surface<false>(S, r, in.t_min, best, nullptr);
  TraceHit h;
  h.t = best.t; h.a = best.a; h.b = best.b; h.prim = best.prim; h.pad = 0;
  out[i] = h;
}

__global__ void __launch_bounds__(128) k_trace_complete(const __grid_constant__ DScene S, const RtbRay* __restrict__ rays,
                                                        long long n, const TraceHit* __restrict__ closest,
                                                        RtbHit* __restrict__ hits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Ray r = load_ray(rays[i]);
  const TraceHit h = closest[i];
  Hit best;
  best.t = h.t; best.a = h.a; best.b = h.b; best.prim = h.prim;
  RtbHit out;
  complete_hit(S, r, best, out);
  hits[i] = out;
}

#if defined(RTB_TRACE_FUSED)
// Investigation arm (tools/repro_fused.sh): traversal and hit completion in ONE kernel, the form that faulted in round 1.
__global__ void __launch_bounds__(128) k_trace_fused(const __grid_constant__ DScene S, const RtbRay* __restrict__ rays,
                                                     long long n, uint32_t flags, RtbHit* __restrict__ hits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RtbRay in = rays[i];
  const Ray r = load_ray(in);
  Hit best;
  hit_reset(best);
  if (flags & RTB_TRACE_BRUTE_FORCE) closest_surface_brute(S, r, in.t_min, best);
  else if (S.n_surface_prims > 0) closest_surface<false>(S, r, in.t_min, best, nullptr);
  RtbHit out;
  complete_hit(S, r, best, out);
  hits[i] = out;
}
#endif

size_t trace_scratch_bytes(int64_t n) { return (size_t)n * sizeof(TraceHit); }

cudaError_t launch_trace(const DScene& S, const RtbRay* d_rays, int64_t n, uint32_t flags, RtbHit* d_hits,
                         void* d_scratch, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 127) / 128);
#if defined(RTB_TRACE_FUSED)
  k_trace_fused<<<blocks, 128, 0, stream>>>(S, d_rays, n, flags, d_hits);
  return cudaGetLastError();
#endif
  k_trace_closest<<<blocks, 128, 0, stream>>>(S, d_rays, n, flags, static_cast<TraceHit*>(d_scratch));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_trace_complete<<<blocks, 128, 0, stream>>>(S, d_rays, n, static_cast<const TraceHit*>(d_scratch), d_hits);
  return cudaGetLastError();
}

__global__ void k_medium_interval(const __grid_constant__ DScene S, int medium, const RtbRay* __restrict__ rays,
                                  long long n, double* __restrict__ t0, double* __restrict__ t1) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RtbRay in = rays[i];
  Ray r;
  r.ox = in.origin[0]; r.oy = in.origin[1]; r.oz = in.origin[2];
  r.dx = in.direction[0]; r.dy = in.direction[1]; r.dz = in.direction[2];
  r.time = in.time;
  double a, b;
  if (medium_interval(S, S.media[medium], r, a, b)) { t0[i] = a; t1[i] = b; }
  else { t0[i] = t1[i] = RTB_INF - RTB_INF; /* NaN */ }
}
cudaError_t launch_medium_interval(const DScene& S, int medium, const RtbRay* d_rays, int64_t n, double* d_t0,
                                   double* d_t1, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_medium_interval<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(S, medium, d_rays, n, d_t0, d_t1);
  return cudaGetLastError();
}

__global__ void k_eval_texture(const __grid_constant__ DScene S, int texture, const double* __restrict__ uvp,
                               long long n, double* __restrict__ rgb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 c = texture_value(S, texture, (float)uvp[5 * i], (float)uvp[5 * i + 1], uvp[5 * i + 2], uvp[5 * i + 3],
                             uvp[5 * i + 4]);
  rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
}
cudaError_t launch_eval_texture(const DScene& S, int texture, const double* d_uvp, int64_t n, double* d_rgb,
                                cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_eval_texture<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(S, texture, d_uvp, n, d_rgb);
  return cudaGetLastError();
}

__global__ void k_eval_light_pdf(const __grid_constant__ DScene S, const double* __restrict__ od, long long n,
                                 double* __restrict__ pdf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ray probe;
  probe.ox = od[6 * i]; probe.oy = od[6 * i + 1]; probe.oz = od[6 * i + 2];
  probe.dx = od[6 * i + 3]; probe.dy = od[6 * i + 4]; probe.dz = od[6 * i + 5];
  probe.time = 0.;
  double sum = 0.;
  RTB_LOOP_ENTER();
  for (int k = 0; k < S.n_lights; k++) sum += light_pdf_one(S.lights[k], probe);
  RTB_LOOP_LEAVE();
  pdf[i] = sum * (1. / (double)S.n_lights);
}
cudaError_t launch_eval_light_pdf(const DScene& S, const double* d_od, int64_t n, double* d_pdf, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_eval_light_pdf<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(S, d_od, n, d_pdf);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// output stage: write_color  reference src/color.rs:8-33 (sRGB OETF :53-59, clamp [0,0.999],
// `(256*x) as u8` saturating, NaN -> 0)
// ------------------------------------------------------------------------------------------------
__global__ void k_write_color(const double* __restrict__ pixels, long long n, double scale, double exposure,
                              uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = pixels[i] * scale;
  if (exposure > 0.) x = 1. - pow(2.718281828459045, -exposure * x);
  x = (x <= 0.0031308) ? 12.92 * x : 1.055 * pow(x, 1. / 2.4) - 0.055;
  const double c = x < 0. ? 0. : (x > 0.999 ? 0.999 : x);
  const double y = 256. * c;
  out[i] = (y == y && y > 0.) ? (uint8_t)(y >= 255. ? 255 : (int)y) : 0;
}
cudaError_t launch_write_color(const double* d_pixels, int64_t n_values, double spp, double exposure, uint8_t* d_out,
                               cudaStream_t stream) {
  if (n_values <= 0) return cudaSuccess;
  k_write_color<<<(unsigned)((n_values + 255) / 256), 256, 0, stream>>>(d_pixels, n_values, 1.0 / spp, exposure, d_out);
  return cudaGetLastError();
}

__global__ void k_accum_to_f64(const unsigned long long* __restrict__ accum, long long n, double* __restrict__ rgb, int add) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool poisoned = accum[4 * i + 3] >= ACCUM_POISON;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double v = (double)(long long)accum[4 * i + c] * (1.0 / ACCUM_SCALE);
    if (poisoned) v = RTB_INF - RTB_INF;
    rgb[3 * i + c] = add ? rgb[3 * i + c] + v : v;
  }
}
cudaError_t launch_accum_to_f64(const unsigned long long* d_accum, int64_t n_pixels, double* d_pixels_rgb, int add, cudaStream_t stream) {
  if (n_pixels <= 0) return cudaSuccess;
  k_accum_to_f64<<<(unsigned)((n_pixels + 255) / 256), 256, 0, stream>>>(d_accum, n_pixels, d_pixels_rgb, add);
  return cudaGetLastError();
}

// Dielectric::scatter known-answer hook: in = n x {d[3], normal[3], front, ir, u}, out = n x direction[3]
__global__ void k_eval_dielectric(const double* __restrict__ in, long long n, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* a = in + 9 * i;
  const V3 d = dielectric_direction(v3((float)a[0], (float)a[1], (float)a[2]), v3((float)a[3], (float)a[4], (float)a[5]), a[6] != 0., (float)a[7], (float)a[8]);
  out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
}
cudaError_t launch_eval_dielectric(const double* d_in, int64_t n, double* d_out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_eval_dielectric<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_in, n, d_out);
  return cudaGetLastError();
}

// Random123 known-answer hook: the device's own philox4x32_10 (rtb_device.cuh) on caller-given counters / keys
__global__ void k_philox(const uint32_t* __restrict__ ctr_key, long long n, uint32_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t o[4];
  philox4x32_10(ctr_key[6 * i], ctr_key[6 * i + 1], ctr_key[6 * i + 2], ctr_key[6 * i + 3], ctr_key[6 * i + 4], ctr_key[6 * i + 5], o);
  out[4 * i] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2]; out[4 * i + 3] = o[3];
}
cudaError_t launch_philox(const uint32_t* d_ctr_key, int64_t n, uint32_t* d_out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_philox<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_ctr_key, n, d_out);
  return cudaGetLastError();
}

}  // namespace rtb
