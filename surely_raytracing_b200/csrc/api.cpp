// api.cpp -- the C ABI of include/rtb200.h: scene upload and kernel orchestration.
// There is deliberately no CPU path here: without a CUDA device every compute entry point
// returns RTB_ERR_NO_DEVICE / RTB_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "device_scene.h"
#include "flatten.h"
#include "kernels.h"

using namespace rtb;

namespace {

thread_local std::string g_err;

int set_err(int code, const std::string& m) { g_err = m; return code; }
int cuda_err(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return RTB_ERR_CUDA;
}
#define CU(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return cuda_err(e__, #call); \
  } while (0)

// Process-wide cache of the large buffers (wavefront workspace, pinned download staging), keyed by
// device: render_par-style callers create and destroy a scene per image, and re-allocating ~0.7 GB
// for each call would dominate short renders.  Freed at process exit.
struct BigBufferCache {
  std::mutex mu;
  struct Entry { int device; int tag; void* p; size_t bytes; bool pinned; bool in_use; };
  std::vector<Entry> entries;
  void* acquire(int device, int tag, size_t bytes, bool pinned, size_t* got, cudaError_t* err) {
    std::lock_guard<std::mutex> lk(mu);
    *err = cudaSuccess;
    for (Entry& e : entries)
      if (!e.in_use && e.device == device && e.tag == tag && e.pinned == pinned && e.bytes >= bytes) { e.in_use = true; *got = e.bytes; return e.p; }
    for (size_t i = 0; i < entries.size(); i++)  // drop an idle, too-small buffer of the same kind
      if (!entries[i].in_use && entries[i].device == device && entries[i].tag == tag && entries[i].pinned == pinned) {
        if (pinned) cudaFreeHost(entries[i].p); else cudaFree(entries[i].p);
        entries.erase(entries.begin() + i);
        break;
      }
    void* p = nullptr;
    *err = pinned ? cudaMallocHost(&p, bytes) : cudaMalloc(&p, bytes);
    if (*err != cudaSuccess) return nullptr;
    entries.push_back(Entry{device, tag, p, bytes, pinned, true});
    *got = bytes;
    return p;
  }
  void release(void* p) {
    std::lock_guard<std::mutex> lk(mu);
    for (Entry& e : entries)
      if (e.p == p) e.in_use = false;
  }
};
BigBufferCache& big_cache() { static BigBufferCache c; return c; }

struct CachedBuffer {  // a lease on a BigBufferCache entry
  void* p = nullptr;
  size_t bytes = 0;
  // `tag` keeps buffer classes apart (a freed 0.7 GB workspace must not be leased as a 2 MB scene block)
  cudaError_t reserve(int device, int tag, size_t n, bool pinned) {
    if (n <= bytes) return cudaSuccess;
    if (p) big_cache().release(p);
    p = nullptr; bytes = 0;
    cudaError_t e;
    p = big_cache().acquire(device, tag, n, pinned, &bytes, &e);
    return e;
  }
  ~CachedBuffer() { if (p) big_cache().release(p); }
};

struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  ~DeviceBuffer() { if (p) cudaFree(p); }
};

}  // namespace

struct rtb_scene {
  int device = 0;
  HostScene host;
  DScene dev{};
  CachedBuffer accum;      // float4[w*h]
  DeviceBuffer stats;      // DStats
  DeviceBuffer scratch_a;  // harness inputs
  DeviceBuffer scratch_b;  // harness outputs
  DeviceBuffer scratch_c;
  CachedBuffer scene_dev;  // all scene arrays, one block
  CachedBuffer scene_host; // pinned staging of that block
  size_t upload_bytes = 0;
  CachedBuffer workspace;  // wavefront queues (leased from the process-wide cache)
  CachedBuffer staging;    // pinned host staging of the accumulation buffer
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  WavefrontContext wf{};
  bool wf_ready = false;
  RtbStats last{};
  bool stats_pending = false;
  ~rtb_scene() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (wf_ready) wavefront_context_destroy(&wf);
  }
};

extern "C" {

int rtb_version(void) { return RTB_ABI_VERSION; }

int rtb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* rtb_last_error(void) { return g_err.c_str(); }

int rtb_scene_create(const RtbSceneDesc* desc, int device, rtb_scene** out) {
  if (!desc || !out) return set_err(RTB_ERR_INVALID, "null argument");
  *out = nullptr;
  rtb_scene* s = new rtb_scene();
  std::string err;
  int rc = flatten_scene(*desc, s->host, err);
  if (rc != RTB_OK) { delete s; return set_err(rc, err); }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    delete s;
    return set_err(RTB_ERR_NO_DEVICE, "no CUDA device visible (this backend has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) { delete s; return set_err(RTB_ERR_INVALID, "device index out of range"); }
  s->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete s; return cuda_err(e, "cudaSetDevice"); }
  const HostScene& h = s->host;
  DScene& D = s->dev;
  // One packed upload: every scene array goes into ONE pinned staging block and ONE device block
  // (both leased from the process-wide cache), copied with a single cudaMemcpyAsync.
  {
    struct Part { const void* src; size_t bytes; size_t off; };
    Part parts[14] = {
        {h.nodes.data(), h.nodes.size() * sizeof(float4), 0},
        {h.prims.data(), h.prims.size() * sizeof(double), 0},
        {h.prim_info.data(), h.prim_info.size() * sizeof(int4), 0},
        {h.xforms.data(), h.xforms.size() * sizeof(double2), 0},
        {h.media.data(), h.media.size() * sizeof(DMedium), 0},
        {h.materials.data(), h.materials.size() * sizeof(DMaterial), 0},
        {h.textures.data(), h.textures.size() * sizeof(DTexture), 0},
        {h.texels.data(), h.texels.size(), 0},
        {h.perlin_vec.data(), h.perlin_vec.size() * sizeof(float4), 0},
        {h.perlin_perm.data(), h.perlin_perm.size(), 0},
        {h.lights.data(), h.lights.size() * sizeof(DLight), 0},
        {h.qnodes.data(), h.qnodes.size() * sizeof(uint4), 0},
        {h.nodes4.data(), h.nodes4.size() * sizeof(float4), 0},
        {h.pre.data(), h.pre.size() * sizeof(DPre), 0},
    };
    size_t total = 0;
    for (Part& p : parts) { p.off = total; total += (std::max<size_t>(p.bytes, 16) + 255) & ~(size_t)255; }
    if ((e = s->scene_dev.reserve(device, 0, total, false)) != cudaSuccess || (e = s->scene_host.reserve(device, 1, total, true)) != cudaSuccess) {
      delete s;
      return cuda_err(e, "scene buffers");
    }
    char* hp = static_cast<char*>(s->scene_host.p);
    for (const Part& p : parts)
      if (p.bytes) std::memcpy(hp + p.off, p.src, p.bytes);
    if ((e = cudaMemcpyAsync(s->scene_dev.p, hp, total, cudaMemcpyHostToDevice, 0)) != cudaSuccess || (e = cudaStreamSynchronize(0)) != cudaSuccess) {
      delete s;
      return cuda_err(e, "scene upload");
    }
    const char* dp = static_cast<const char*>(s->scene_dev.p);
    D.nodes = reinterpret_cast<const float4*>(dp + parts[0].off);
    D.prims = reinterpret_cast<const double2*>(dp + parts[1].off);
    D.prim_info = reinterpret_cast<const int4*>(dp + parts[2].off);
    D.xforms = reinterpret_cast<const double2*>(dp + parts[3].off);
    D.media = reinterpret_cast<const DMedium*>(dp + parts[4].off);
    D.materials = reinterpret_cast<const DMaterial*>(dp + parts[5].off);
    D.textures = reinterpret_cast<const DTexture*>(dp + parts[6].off);
    D.texels = reinterpret_cast<const uint8_t*>(dp + parts[7].off);
    D.perlin_vec = reinterpret_cast<const float4*>(dp + parts[8].off);
    D.perlin_perm = reinterpret_cast<const uint8_t*>(dp + parts[9].off);
    D.lights = reinterpret_cast<const DLight*>(dp + parts[10].off);
    D.qnodes = reinterpret_cast<const uint4*>(dp + parts[11].off);
    D.nodes4 = reinterpret_cast<const float4*>(dp + parts[12].off);
    D.pre = reinterpret_cast<const DPre*>(dp + parts[13].off);
    s->upload_bytes = total;
  }
  D.n_nodes = (int)h.nodes.size() / 4;
  D.n_surface_prims = h.n_surface_prims;
  D.n_prims = (int)h.prim_info.size();
  D.n_media = (int)h.media.size();
  D.n_lights = (int)h.lights.size();
  for (int a = 0; a < 3; a++) { D.grid_base[a] = h.grid_base[a]; D.grid_inv_cell[a] = h.grid_inv_cell[a]; D.grid_cell[a] = h.grid_cell[a]; }
  D.use_qnodes = h.use_qnodes;
  D.use_bvh4 = h.use_bvh4;
  D.spec_bits = h.spec_bits;
  D.scene_mag = h.scene_mag;
  D.multi_leaf = h.multi_leaf;
  D.defer_ok = h.defer_ok;
  D.n_materials = (int)h.materials.size();
  D.n_textures = (int)h.textures.size();
  D.bvh_depth = h.bvh_depth;
  D.flags = h.flags;
  D.seed_lo = (uint32_t)h.seed;
  D.seed_hi = (uint32_t)(h.seed >> 32);
  D.cam = h.cam;
  if ((e = cudaEventCreate(&s->ev0)) != cudaSuccess || (e = cudaEventCreate(&s->ev1)) != cudaSuccess) {
    delete s;
    return cuda_err(e, "cudaEventCreate");
  }
  if ((e = s->stats.reserve(sizeof(DStats))) != cudaSuccess) { delete s; return cuda_err(e, "cudaMalloc stats"); }
  *out = s;
  return RTB_OK;
}

void rtb_scene_destroy(rtb_scene* scene) {
  if (!scene) return;
  cudaSetDevice(scene->device);
  delete scene;
}

int rtb_scene_info(const rtb_scene* s, RtbSceneInfo* info) {
  if (!s || !info) return set_err(RTB_ERR_INVALID, "null argument");
  std::memset(info, 0, sizeof(*info));
  info->image_width = s->host.cam.width;
  info->image_height = s->host.cam.height;
  info->spp_used = s->host.cam.spp;
  info->sqrt_spp = s->host.cam.sqrt_spp;
  info->max_depth = s->host.cam.max_depth;
  info->n_surface_prims = s->host.n_surface_prims;
  info->n_boundary_prims = (int)s->host.prim_info.size() - s->host.n_surface_prims;
  info->n_media = (int)s->host.media.size();
  info->n_bvh_nodes = (int)s->host.nodes.size() / 4;
  info->n_lights = (int)s->host.lights.size();
  info->bvh_depth = s->host.bvh_depth;
  info->device = s->device;
  return RTB_OK;
}

static int check_range(const rtb_scene* s, const RtbRenderParams* p) {
  if (!s || !p) return set_err(RTB_ERR_INVALID, "null argument");
  if (p->sample_begin < 0 || p->sample_end > s->host.cam.spp || p->sample_begin > p->sample_end)
    return set_err(RTB_ERR_INVALID, "sample range outside [0, spp_used]");
  return RTB_OK;
}

// Path slots in flight of the wavefront pipeline (RTB_WF_CAPACITY overrides, for tuning runs).
// Every iteration pays ~55 us of launch gaps and kernel tails whatever the queue size, and the decaying
// tail of a call (no new paths left to start) costs in proportion to it: measured on c4, 1/16 of the
// call's paths is the sweet spot (64 M paths: 4 M slots; 512 M and more: 32 M slots = 4.6 GB of queues).
static int64_t wavefront_capacity(int64_t total_paths) {
  if (const char* e = getenv("RTB_WF_CAPACITY")) {  // read per call: tests shrink the queue to exercise refill and tail
    const long long v = atoll(e);
    if (v >= 1024) return (int64_t)v;
  }
  int64_t cap = 1 << 20;
  while (cap < (1 << 25) && cap * 16 < total_paths) cap <<= 1;
  return cap;
}

static int render_into(rtb_scene* s, const RtbRenderParams* p, float4* d_accum, cudaStream_t stream) {
  CU(cudaSetDevice(s->device));
  const bool collect = p->collect_stats != 0;
  DStats* d_stats = static_cast<DStats*>(s->stats.p);
  CU(cudaMemsetAsync(d_stats, 0, sizeof(DStats), stream));
  int launches = 0;
  CU(cudaEventRecord(s->ev0, stream));
  if (p->sample_end > p->sample_begin) {
    if (p->pipeline != RTB_PIPELINE_MEGAKERNEL) {  // default = wavefront
      const int64_t cap = wavefront_capacity((int64_t)(p->sample_end - p->sample_begin) * s->host.cam.width * s->host.cam.height);
      const size_t ws = wavefront_workspace_bytes(s->dev, cap);
      CU(s->workspace.reserve(s->device, 2, ws, false));
      if (!s->wf_ready) {
        CU(wavefront_context_create(&s->wf));
        s->wf_ready = true;
      }
      CU(launch_render_wavefront(s->dev, s->wf, p->sample_begin, p->sample_end, d_accum, d_stats, collect, s->workspace.p,
                                 s->workspace.bytes, cap, stream, &launches));
    } else {
      CU(launch_render_mega(s->dev, p->sample_begin, p->sample_end, d_accum, d_stats, collect, stream, &launches));
    }
  }
  CU(cudaEventRecord(s->ev1, stream));
  s->last = RtbStats{};
  s->last.kernel_launches = (uint64_t)launches;
  s->last.paths = (uint64_t)(p->sample_end - p->sample_begin) * (uint64_t)s->host.cam.width * (uint64_t)s->host.cam.height;
  s->stats_pending = true;
  return RTB_OK;
}

int rtb_render_stats(rtb_scene* s, RtbStats* stats) {
  if (!s || !stats) return set_err(RTB_ERR_INVALID, "null argument");
  CU(cudaSetDevice(s->device));
  if (s->stats_pending) {
    CU(cudaEventSynchronize(s->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    DStats h{};
    CU(cudaMemcpy(&h, s->stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    s->last.device_ms = ms;
    s->last.segments = h.segments;
    s->last.node_visits = h.node_visits;
    s->last.prim_tests = h.prim_tests;
    s->last.medium_probes = h.medium_probes;
    s->last.nonfinite_samples = h.nonfinite;
    s->stats_pending = false;
  }
  *stats = s->last;
  return RTB_OK;
}

int rtb_render_device(rtb_scene* s, const RtbRenderParams* p, void* d_accum_rgba, void* cuda_stream) {
  int rc = check_range(s, p);
  if (rc != RTB_OK) return rc;
  if (!d_accum_rgba) return set_err(RTB_ERR_INVALID, "null accumulation buffer");
  return render_into(s, p, static_cast<float4*>(d_accum_rgba), static_cast<cudaStream_t>(cuda_stream));
}

int rtb_render(rtb_scene* s, const RtbRenderParams* p, double* pixels_rgb, RtbStats* stats) {
  int rc = check_range(s, p);
  if (rc != RTB_OK) return rc;
  if (!pixels_rgb) return set_err(RTB_ERR_INVALID, "null pixel buffer");
  CU(cudaSetDevice(s->device));
  const size_t n = (size_t)s->host.cam.width * s->host.cam.height;
  CU(s->accum.reserve(s->device, 4, n * sizeof(float4), false));
  CU(cudaMemsetAsync(s->accum.p, 0, n * sizeof(float4), 0));
  rc = render_into(s, p, static_cast<float4*>(s->accum.p), 0);
  if (rc != RTB_OK) return rc;
  CU(s->staging.reserve(s->device, 3, n * sizeof(float4), true));
  const float4* h = static_cast<const float4*>(s->staging.p);
  CU(cudaMemcpyAsync(s->staging.p, s->accum.p, n * sizeof(float4), cudaMemcpyDeviceToHost, 0));
  CU(cudaStreamSynchronize(0));
  for (size_t i = 0; i < n; i++) {  // `row[i] = row[i] + color` (Q24): accumulate INTO the caller's sums
    pixels_rgb[3 * i + 0] += (double)h[i].x;
    pixels_rgb[3 * i + 1] += (double)h[i].y;
    pixels_rgb[3 * i + 2] += (double)h[i].z;
  }
  RtbStats tmp;
  rc = rtb_render_stats(s, &tmp);
  if (rc != RTB_OK) return rc;
  if (stats) *stats = tmp;
  return RTB_OK;
}

int rtb_camera_rays(const rtb_scene* s, RtbRay* rays) {
  if (!s || !rays) return set_err(RTB_ERR_INVALID, "null argument");
  const DCamera& c = s->host.cam;
  for (int j = 0; j < c.height; j++)
    for (int i = 0; i < c.width; i++) {
      RtbRay& r = rays[(size_t)j * c.width + i];
      for (int a = 0; a < 3; a++) {
        const double pc = c.pixel00[a] + ((double)i * c.du[a]) + ((double)j * c.dv[a]);
        r.origin[a] = c.center[a];
        r.direction[a] = pc - c.center[a];
      }
      r.time = 0.;
      r.t_min = 0.0001;
    }
  return RTB_OK;
}

int rtb_trace(rtb_scene* s, const RtbRay* rays, int64_t n, uint32_t flags, RtbHit* hits) {
  if (!s || (n > 0 && (!rays || !hits)) || n < 0) return set_err(RTB_ERR_INVALID, "bad argument");
  if (n == 0) return RTB_OK;
  CU(cudaSetDevice(s->device));
  CU(s->scratch_a.reserve((size_t)n * sizeof(RtbRay)));
  CU(s->scratch_b.reserve((size_t)n * sizeof(RtbHit)));
  CU(s->scratch_c.reserve(trace_scratch_bytes(n)));
  CU(cudaMemcpy(s->scratch_a.p, rays, (size_t)n * sizeof(RtbRay), cudaMemcpyHostToDevice));
  CU(launch_trace(s->dev, static_cast<const RtbRay*>(s->scratch_a.p), n, flags, static_cast<RtbHit*>(s->scratch_b.p), s->scratch_c.p, 0));
  CU(cudaMemcpy(hits, s->scratch_b.p, (size_t)n * sizeof(RtbHit), cudaMemcpyDeviceToHost));
  return RTB_OK;
}

int rtb_medium_interval(rtb_scene* s, int32_t medium, const RtbRay* rays, int64_t n, double* t_enter, double* t_exit) {
  if (!s || n < 0 || (n > 0 && (!rays || !t_enter || !t_exit))) return set_err(RTB_ERR_INVALID, "bad argument");
  if (medium < 0 || medium >= (int)s->host.media.size()) return set_err(RTB_ERR_INVALID, "no such medium");
  if (n == 0) return RTB_OK;
  CU(cudaSetDevice(s->device));
  CU(s->scratch_a.reserve((size_t)n * sizeof(RtbRay)));
  CU(s->scratch_b.reserve((size_t)n * sizeof(double)));
  CU(s->scratch_c.reserve((size_t)n * sizeof(double)));
  CU(cudaMemcpy(s->scratch_a.p, rays, (size_t)n * sizeof(RtbRay), cudaMemcpyHostToDevice));
  CU(launch_medium_interval(s->dev, medium, static_cast<const RtbRay*>(s->scratch_a.p), n, static_cast<double*>(s->scratch_b.p),
                            static_cast<double*>(s->scratch_c.p), 0));
  CU(cudaMemcpy(t_enter, s->scratch_b.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(t_exit, s->scratch_c.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return RTB_OK;
}

int rtb_eval_texture(rtb_scene* s, int32_t texture, const double* uvp, int64_t n, double* rgb_out) {
  if (!s || n < 0 || (n > 0 && (!uvp || !rgb_out))) return set_err(RTB_ERR_INVALID, "bad argument");
  if (texture < 0 || texture >= (int)s->host.textures.size()) return set_err(RTB_ERR_INVALID, "no such texture");
  if (n == 0) return RTB_OK;
  CU(cudaSetDevice(s->device));
  CU(s->scratch_a.reserve((size_t)n * 5 * sizeof(double)));
  CU(s->scratch_b.reserve((size_t)n * 3 * sizeof(double)));
  CU(cudaMemcpy(s->scratch_a.p, uvp, (size_t)n * 5 * sizeof(double), cudaMemcpyHostToDevice));
  CU(launch_eval_texture(s->dev, texture, static_cast<const double*>(s->scratch_a.p), n, static_cast<double*>(s->scratch_b.p), 0));
  CU(cudaMemcpy(rgb_out, s->scratch_b.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  return RTB_OK;
}

int rtb_eval_light_pdf(rtb_scene* s, const double* origin_dir, int64_t n, double* pdf_out) {
  if (!s || n < 0 || (n > 0 && (!origin_dir || !pdf_out))) return set_err(RTB_ERR_INVALID, "bad argument");
  if (s->host.lights.empty()) return set_err(RTB_ERR_INVALID, "scene has no lights");
  if (n == 0) return RTB_OK;
  CU(cudaSetDevice(s->device));
  CU(s->scratch_a.reserve((size_t)n * 6 * sizeof(double)));
  CU(s->scratch_b.reserve((size_t)n * sizeof(double)));
  CU(cudaMemcpy(s->scratch_a.p, origin_dir, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice));
  CU(launch_eval_light_pdf(s->dev, static_cast<const double*>(s->scratch_a.p), n, static_cast<double*>(s->scratch_b.p), 0));
  CU(cudaMemcpy(pdf_out, s->scratch_b.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return RTB_OK;
}

int rtb_write_color(rtb_scene* s, const double* pixels_rgb, int64_t n_pixels, double spp, double exposure, uint8_t* rgb8_out) {
  if (!s || n_pixels < 0 || (n_pixels > 0 && (!pixels_rgb || !rgb8_out))) return set_err(RTB_ERR_INVALID, "bad argument");
  if (n_pixels == 0) return RTB_OK;
  CU(cudaSetDevice(s->device));
  const size_t nv = (size_t)n_pixels * 3;
  CU(s->scratch_a.reserve(nv * sizeof(double)));
  CU(s->scratch_b.reserve(nv));
  CU(cudaMemcpy(s->scratch_a.p, pixels_rgb, nv * sizeof(double), cudaMemcpyHostToDevice));
  CU(launch_write_color(static_cast<const double*>(s->scratch_a.p), (int64_t)nv, spp, exposure, static_cast<uint8_t*>(s->scratch_b.p), 0));
  CU(cudaMemcpy(rgb8_out, s->scratch_b.p, nv, cudaMemcpyDeviceToHost));
  return RTB_OK;
}

}  // extern "C"
