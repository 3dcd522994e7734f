// api.cpp -- the C ABI of include/rtb200.h: scene upload and kernel orchestration.
// There is deliberately no CPU path here: without a CUDA device every compute entry point
// returns RTB_ERR_NO_DEVICE / RTB_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
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
  cudaGetLastError();  // do not leave the error for the next, unrelated call
  return RTB_ERR_CUDA;
}
#define CU(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return cuda_err(e__, #call); \
  } while (0)

// every extern "C" body runs inside this: no C++ exception crosses the C boundary
template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return set_err(RTB_ERR_INVALID, "out of host memory");
  } catch (const std::exception& e) {
    return set_err(RTB_ERR_INVALID, std::string("internal error: ") + e.what());
  } catch (...) {
    return set_err(RTB_ERR_INVALID, "internal error");
  }
}

// Process-wide cache of the large buffers (wavefront workspace, pinned staging), keyed by device:
// render_par-style callers create and destroy a scene per image, and re-allocating ~0.7 GB for each call
// would dominate short renders.  rtb_trim_cache() releases what is idle; the rest goes at process exit.
struct BigBufferCache {
  std::mutex mu;
  struct Entry { int device; int tag; void* p; size_t bytes; bool pinned; bool in_use; };
  std::vector<Entry> entries;
  static void free_entry(const Entry& e) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(e.device);
    if (e.pinned) cudaFreeHost(e.p); else cudaFree(e.p);
    cudaSetDevice(cur);
  }
  void* acquire(int device, int tag, size_t bytes, bool pinned, size_t* got, cudaError_t* err) {
    std::lock_guard<std::mutex> lk(mu);
    *err = cudaSuccess;
    for (Entry& e : entries)
      if (!e.in_use && e.device == device && e.tag == tag && e.pinned == pinned && e.bytes >= bytes) { e.in_use = true; *got = e.bytes; return e.p; }
    for (size_t i = 0; i < entries.size(); i++)  // drop an idle, too-small buffer of the same kind
      if (!entries[i].in_use && entries[i].device == device && entries[i].tag == tag && entries[i].pinned == pinned) {
        free_entry(entries[i]);
        entries.erase(entries.begin() + i);
        break;
      }
    void* p = nullptr;
    *err = pinned ? cudaMallocHost(&p, bytes) : cudaMalloc(&p, bytes);
    if (*err != cudaSuccess) { cudaGetLastError(); return nullptr; }
    entries.push_back(Entry{device, tag, p, bytes, pinned, true});
    *got = bytes;
    return p;
  }
  void release(void* p) {
    std::lock_guard<std::mutex> lk(mu);
    for (Entry& e : entries)
      if (e.p == p) e.in_use = false;
  }
  int64_t trim() {
    std::lock_guard<std::mutex> lk(mu);
    int64_t freed = 0;
    for (size_t i = 0; i < entries.size();) {
      if (!entries[i].in_use) {
        free_entry(entries[i]);
        freed += (int64_t)entries[i].bytes;
        entries.erase(entries.begin() + i);
      } else {
        i++;
      }
    }
    return freed;
  }
};
BigBufferCache& big_cache() { static BigBufferCache c; return c; }

struct CachedBuffer {  // a lease on a BigBufferCache entry
  void* p = nullptr;
  size_t bytes = 0;
  // `tag` keeps buffer classes apart (a freed 0.7 GB workspace must not be leased as a 2 MB scene block)
  cudaError_t reserve(int device, int tag, size_t n, bool pinned) {
    if (n <= bytes) return cudaSuccess;
    drop();
    cudaError_t e;
    p = big_cache().acquire(device, tag, n, pinned, &bytes, &e);
    return e;
  }
  void drop() {
    if (p) big_cache().release(p);
    p = nullptr; bytes = 0;
  }
  ~CachedBuffer() { drop(); }
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

constexpr int64_t kMegaBelowDefault = 1 << 19;  // paths per call under which the megakernel wins (the wavefront is launch-bound there)

}  // namespace

struct rtb_scene {
  int device = 0;
  std::shared_ptr<const HostScene> host;
  DScene dev{};
  CachedBuffer accum;      // uint64 x 4 per pixel
  CachedBuffer stats;      // DStats
  CachedBuffer counters;   // pinned mirror of the wavefront counters
  DeviceBuffer scratch_a;  // harness inputs
  DeviceBuffer scratch_b;  // harness outputs
  DeviceBuffer scratch_c;
  CachedBuffer scene_dev;  // all scene arrays, one block
  CachedBuffer scene_host; // pinned staging of that block
  size_t upload_bytes = 0;
  CachedBuffer workspace;  // wavefront queues (leased from the process-wide cache)
  CachedBuffer staging;    // pinned host staging of the f64 sums
  CachedBuffer sums;       // device f64 sums on their way to the host (cached: a cudaFree per scene costs a device sync)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  WavefrontContext wf{};
  bool wf_ready = false;
  WavefrontOptions opt{};
  int64_t mega_below = -1;
  RtbStats last{};
  bool stats_pending = false;
  ~rtb_scene() {
    // a caller's stream may still run kernels of this scene (rtb_render_device does not synchronise): wait for
    // the last recorded work before the leased blocks go back to the cache, where another scene may take them
    if (ev1) { cudaEventSynchronize(ev1); cudaEventDestroy(ev1); }
    if (ev0) cudaEventDestroy(ev0);
    cudaGetLastError();
  }
};

namespace {

// upload the flattened scene to `device` and fill the device-side description
int scene_upload(rtb_scene* s, int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(RTB_ERR_NO_DEVICE, "no CUDA device visible (this backend has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return set_err(RTB_ERR_INVALID, "device index out of range");
  s->device = device;
  CU(cudaSetDevice(device));
  const HostScene& h = *s->host;
  DScene& D = s->dev;
  // One packed upload: every scene array goes into ONE pinned staging block and ONE device block
  // (both leased from the process-wide cache), copied with a single cudaMemcpyAsync.
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
  CU(s->scene_dev.reserve(device, 0, total, false));
  CU(s->scene_host.reserve(device, 1, total, true));
  char* hp = static_cast<char*>(s->scene_host.p);
  for (const Part& p : parts)
    if (p.bytes) std::memcpy(hp + p.off, p.src, p.bytes);
  // cudaStreamPerThread: the upload must not order against (or behind) other scenes' work on the legacy stream
  CU(cudaMemcpyAsync(s->scene_dev.p, hp, total, cudaMemcpyHostToDevice, cudaStreamPerThread));
  CU(cudaStreamSynchronize(cudaStreamPerThread));
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
  D.n_suns = h.n_suns;
  for (int k = 0; k < h.n_suns; k++) D.suns[k] = h.suns[k];
  CU(cudaEventCreate(&s->ev0));
  CU(cudaEventCreate(&s->ev1));
  CU(s->stats.reserve(device, 6, 256, false));
  return RTB_OK;
}

// per-scene host state of the wavefront driver; the small pinned / device blocks come from the process-wide cache
// (a cudaMallocHost per scene was a millisecond of every render_par-style call)
int wavefront_ready(rtb_scene* s) {
  if (s->wf_ready) return RTB_OK;
  CU(s->counters.reserve(s->device, 5, std::max<size_t>(256, wavefront_counters_bytes()), true));
  s->wf.host_counters = s->counters.p;
  s->wf.sms = 148;
  CU(cudaDeviceGetAttribute(&s->wf.sms, cudaDevAttrMultiProcessorCount, s->device));
  s->wf_ready = true;
  return RTB_OK;
}

int check_range(const rtb_scene* s, const RtbRenderParams* p) {
  if (!s || !p) return set_err(RTB_ERR_INVALID, "null argument");
  if (p->sample_begin < 0 || p->sample_end > s->host->cam.spp || p->sample_begin > p->sample_end)
    return set_err(RTB_ERR_INVALID, "sample range outside [0, spp_used]");
  if (p->pipeline < RTB_PIPELINE_DEFAULT || p->pipeline > RTB_PIPELINE_WAVEFRONT) return set_err(RTB_ERR_INVALID, "unknown pipeline");
  return RTB_OK;
}

// Path slots in flight of the wavefront pipeline (RTB_OPT_WF_CAPACITY overrides, for tuning runs and tests).
// Every iteration pays ~50 us of launch gaps and kernel tails whatever the queue size, and the decaying
// tail of a call (no new paths left to start) costs in proportion to it: measured on c4, 1/16 of the
// call's paths is the sweet spot (64 M paths: 4 M slots; 512 M and more: 32 M slots = 5.1 GB of queues).
int64_t wavefront_capacity(const rtb_scene* s, int64_t total_paths) {
  if (s->opt.capacity >= 1024) return s->opt.capacity;
  // ... and a small call is launch-bound: c1 (4.4 M paths) takes 3.00 ms with 1 M slots, 2.60 with 2 M, 2.43 with all of
  // its paths in flight at once -- so never fewer than 4 M slots, or the whole call when it is smaller than that.
  int64_t cap = 1 << 22;
  while (cap < (1 << 25) && cap * 16 < total_paths) cap <<= 1;
  return std::max<int64_t>(1024, std::min<int64_t>(cap, (total_paths + 255) & ~(int64_t)255));
}

int render_into(rtb_scene* s, const RtbRenderParams* p, unsigned long long* d_accum, cudaStream_t stream) {
  CU(cudaSetDevice(s->device));
  const bool collect = p->collect_stats != 0;
  DStats* d_stats = static_cast<DStats*>(s->stats.p);
  CU(cudaMemsetAsync(d_stats, 0, sizeof(DStats), stream));
  int launches = 0;
  double stage_ms[3] = {0., 0., 0.};
  CU(cudaEventRecord(s->ev0, stream));
  const int64_t n_strata = p->sample_end - p->sample_begin;
  const int64_t n_pixels = (int64_t)s->host->cam.width * s->host->cam.height;
  const int64_t total_paths = n_strata * n_pixels;
  if (n_strata > 0 && s->host->cam.max_depth <= 0) {
    // ray_color returns (0,0,0) at depth <= 0 (src/render.rs:260-262): nothing to trace, the samples still count
    CU(launch_add_count(d_accum, n_pixels, (unsigned long long)n_strata, stream));
    launches++;
  } else if (n_strata > 0) {
    const int64_t mega_below = s->mega_below >= 0 ? s->mega_below : kMegaBelowDefault;
    const bool mega = p->pipeline == RTB_PIPELINE_MEGAKERNEL || (p->pipeline == RTB_PIPELINE_DEFAULT && total_paths < mega_below);
    if (!mega) {
      int64_t cap = wavefront_capacity(s, total_paths);
      for (;;) {  // a queue that does not fit is halved (more iterations, same image) down to 64 K slots
        cudaError_t e = s->workspace.reserve(s->device, 2, wavefront_workspace_bytes(s->dev, cap), false);
        if (e == cudaSuccess) break;
        if (e != cudaErrorMemoryAllocation || cap <= (1 << 16)) return cuda_err(e, "wavefront workspace");
        cap >>= 1;
      }
      if (const int wrc = wavefront_ready(s)) return wrc;
      s->opt.defer_rare = s->opt.no_defer_rare ? 0 : 1;
      CU(launch_render_wavefront(s->dev, s->wf, s->opt, p->sample_begin, p->sample_end, d_accum, d_stats, collect, s->workspace.p,
                                 s->workspace.bytes, cap, stream, &launches, stage_ms));
    } else {
      CU(launch_render_mega(s->dev, p->sample_begin, p->sample_end, d_accum, d_stats, collect, stream, &launches));
    }
  }
  CU(cudaEventRecord(s->ev1, stream));
  s->last = RtbStats{};
  s->last.kernel_launches = (uint64_t)launches;
  s->last.paths = (uint64_t)total_paths;
  for (int j = 0; j < 3; j++) s->last.stage_ms[j] = stage_ms[j];
  s->stats_pending = true;
  return RTB_OK;
}

int fetch_stats(rtb_scene* s, RtbStats* stats) {
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
    s->last.exact_tests = h.exact_tests;
    s->last.overflow_rays = h.overflows;
    s->stats_pending = false;
  }
  *stats = s->last;
  return RTB_OK;
}

// device accumulation buffer -> += host f64 sums (`row[i] = row[i] + color`, Q24)
int accum_to_host(rtb_scene* s, const unsigned long long* d_accum, double* pixels_rgb, cudaStream_t stream) {
  const size_t n = (size_t)s->host->cam.width * s->host->cam.height;
  CU(s->sums.reserve(s->device, 7, n * 3 * sizeof(double), false));
  CU(s->staging.reserve(s->device, 3, n * 3 * sizeof(double), true));
  CU(launch_accum_to_f64(d_accum, (int64_t)n, static_cast<double*>(s->sums.p), 0, stream));
  CU(cudaMemcpyAsync(s->staging.p, s->sums.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  const double* h = static_cast<const double*>(s->staging.p);
  for (size_t i = 0; i < 3 * n; i++) pixels_rgb[i] += h[i];
  return RTB_OK;
}

int scene_create(std::shared_ptr<const HostScene> host, int device, rtb_scene** out) {
  std::unique_ptr<rtb_scene> s(new rtb_scene());
  s->host = std::move(host);
  s->opt.finish_below = -1;
  const int rc = scene_upload(s.get(), device);
  if (rc != RTB_OK) return rc;
  *out = s.release();
  return RTB_OK;
}

// ---- NCCL, resolved at run time: single-GPU users never need the library -----------------------------------
struct Nccl {
  typedef void* comm_t;
  int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
  Nccl() {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { why = "libnccl.so.2 not found (dlopen)"; return; }
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    Reduce = reinterpret_cast<decltype(Reduce)>(dlsym(h, "ncclReduce"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(h, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    ok = CommInitAll && CommDestroy && Reduce && GroupStart && GroupEnd && GetErrorString;
    if (!ok) why = "libnccl.so.2 lacks an expected symbol";
  }
};
Nccl& nccl() { static Nccl n; return n; }
constexpr int kNcclInt64 = 4, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

// communicators are expensive to create (~1 s for 8 GPUs): one set per device list, kept for the process
struct CommCache {
  std::mutex mu;
  std::vector<int> devices;
  std::vector<Nccl::comm_t> comms;
  int get(const std::vector<int>& devs, std::vector<Nccl::comm_t>& out) {
    std::lock_guard<std::mutex> lk(mu);
    if (devs != devices) {
      for (Nccl::comm_t c : comms) nccl().CommDestroy(c);
      comms.assign(devs.size(), nullptr);
      devices.clear();
      const int rc = nccl().CommInitAll(comms.data(), (int)devs.size(), devs.data());
      if (rc != 0) { comms.clear(); return set_err(RTB_ERR_CUDA, std::string("ncclCommInitAll: ") + nccl().GetErrorString(rc)); }
      devices = devs;
    }
    out = comms;
    return RTB_OK;
  }
};
CommCache& comm_cache() { static CommCache c; return c; }

}  // namespace

extern "C" {

int rtb_version(void) { return RTB_ABI_VERSION; }

int rtb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* rtb_last_error(void) { return g_err.c_str(); }

int rtb_scene_create(const RtbSceneDesc* desc, int device, rtb_scene** out) {
  return guarded([&]() -> int {
    if (!desc || !out) return set_err(RTB_ERR_INVALID, "null argument");
    *out = nullptr;
    auto host = std::make_shared<HostScene>();
    std::string err;
    const int rc = flatten_scene(*desc, *host, err);
    if (rc != RTB_OK) return set_err(rc, err);
    return scene_create(std::move(host), device, out);
  });
}

void rtb_scene_destroy(rtb_scene* scene) {
  if (!scene) return;
  cudaSetDevice(scene->device);
  delete scene;
}

int rtb_scene_info(const rtb_scene* s, RtbSceneInfo* info) {
  if (!s || !info) return set_err(RTB_ERR_INVALID, "null argument");
  const HostScene& h = *s->host;
  std::memset(info, 0, sizeof(*info));
  info->image_width = h.cam.width;
  info->image_height = h.cam.height;
  info->spp_used = h.cam.spp;
  info->sqrt_spp = h.cam.sqrt_spp;
  info->max_depth = h.cam.max_depth;
  info->n_surface_prims = h.n_surface_prims;
  info->n_boundary_prims = (int)h.prim_info.size() - h.n_surface_prims;
  info->n_media = (int)h.media.size();
  info->n_bvh_nodes = (int)h.nodes.size() / 4;
  info->n_lights = (int)h.lights.size();
  info->bvh_depth = h.bvh_depth;
  info->device = s->device;
  return RTB_OK;
}

namespace {
struct CheckpointHeader {  // 64 bytes
  char magic[8];           // "RTB200CK"
  int32_t abi, width, height, reserved;
  uint64_t seed;
  uint32_t flags, pad;
  uint64_t spare[3];
};
static_assert(sizeof(CheckpointHeader) == 64, "checkpoint header is 64 bytes");
}  // namespace

int rtb_checkpoint_save(rtb_scene* s, const void* d_accum, const char* path) {
  return guarded([&]() -> int {
    if (!s || !d_accum || !path) return set_err(RTB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    CU(cudaDeviceSynchronize());
    const size_t n = (size_t)s->host->cam.width * s->host->cam.height * 4;
    std::vector<unsigned long long> h(n);
    CU(cudaMemcpy(h.data(), d_accum, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CheckpointHeader hd{};
    std::memcpy(hd.magic, "RTB200CK", 8);
    hd.abi = RTB_ABI_VERSION; hd.width = s->host->cam.width; hd.height = s->host->cam.height;
    hd.seed = s->host->seed; hd.flags = s->host->flags;
    FILE* f = std::fopen(path, "wb");
    if (!f) return set_err(RTB_ERR_INVALID, std::string("cannot open ") + path);
    const bool ok = std::fwrite(&hd, sizeof(hd), 1, f) == 1 && std::fwrite(h.data(), sizeof(unsigned long long), n, f) == n;
    if (std::fclose(f) != 0 || !ok) return set_err(RTB_ERR_INVALID, std::string("short write to ") + path);
    return RTB_OK;
  });
}

int rtb_checkpoint_load(rtb_scene* s, void* d_accum, const char* path) {
  return guarded([&]() -> int {
    if (!s || !d_accum || !path) return set_err(RTB_ERR_INVALID, "null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return set_err(RTB_ERR_INVALID, std::string("cannot open ") + path);
    CheckpointHeader hd{};
    const size_t n = (size_t)s->host->cam.width * s->host->cam.height * 4;
    std::vector<unsigned long long> h(n);
    const bool ok = std::fread(&hd, sizeof(hd), 1, f) == 1 && std::fread(h.data(), sizeof(unsigned long long), n, f) == n;
    std::fclose(f);
    if (!ok || std::memcmp(hd.magic, "RTB200CK", 8) != 0) return set_err(RTB_ERR_INVALID, "not a checkpoint of this image size");
    if (hd.abi != RTB_ABI_VERSION || hd.width != s->host->cam.width || hd.height != s->host->cam.height || hd.seed != s->host->seed)
      return set_err(RTB_ERR_INVALID, "checkpoint belongs to another image (size, seed or ABI differ)");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpy(d_accum, h.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    return RTB_OK;
  });
}

int rtb_scene_set_option(rtb_scene* s, int option, int64_t value) {
  if (!s) return set_err(RTB_ERR_INVALID, "null argument");
  switch (option) {
    case RTB_OPT_WF_CAPACITY:
      if (value != 0 && value < 1024) return set_err(RTB_ERR_INVALID, "queue capacity below 1024");
      s->opt.capacity = value;
      return RTB_OK;
    case RTB_OPT_EXACT_LEAVES: s->opt.exact_leaves = value != 0; return RTB_OK;
    case RTB_OPT_SMEM_TOP: s->opt.smem_top = value != 0; return RTB_OK;
    case RTB_OPT_NO_DEFER_RARE: s->opt.no_defer_rare = value != 0; return RTB_OK;
    case RTB_OPT_EXTEND_BLOCKS: s->opt.extend_blocks_per_sm = (int)std::max<int64_t>(0, std::min<int64_t>(32, value)); return RTB_OK;
    case RTB_OPT_FINISH_BELOW: s->opt.finish_below = (int)std::max<int64_t>(-1, std::min<int64_t>(1 << 24, value)); return RTB_OK;
    case RTB_OPT_PROFILE: s->opt.profile = (int)value; return RTB_OK;
    case RTB_OPT_MEGA_BELOW: s->mega_below = value; return RTB_OK;
    default: return set_err(RTB_ERR_INVALID, "unknown option");
  }
}

int64_t rtb_trim_cache(void) { return big_cache().trim(); }

int rtb_render_stats(rtb_scene* s, RtbStats* stats) {
  return guarded([&]() -> int {
    if (!s || !stats) return set_err(RTB_ERR_INVALID, "null argument");
    return fetch_stats(s, stats);
  });
}

int rtb_render_device(rtb_scene* s, const RtbRenderParams* p, void* d_accum, void* cuda_stream) {
  return guarded([&]() -> int {
    int rc = check_range(s, p);
    if (rc != RTB_OK) return rc;
    if (!d_accum) return set_err(RTB_ERR_INVALID, "null accumulation buffer");
    return render_into(s, p, static_cast<unsigned long long*>(d_accum), static_cast<cudaStream_t>(cuda_stream));
  });
}

int rtb_accum_to_pixels(rtb_scene* s, const void* d_accum, double* pixels_rgb) {
  return guarded([&]() -> int {
    if (!s || !d_accum || !pixels_rgb) return set_err(RTB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    CU(cudaDeviceSynchronize());  // the buffer may have been produced on any stream of the caller
    return accum_to_host(s, static_cast<const unsigned long long*>(d_accum), pixels_rgb, cudaStreamPerThread);
  });
}

int rtb_render(rtb_scene* s, const RtbRenderParams* p, double* pixels_rgb, RtbStats* stats) {
  return guarded([&]() -> int {
    int rc = check_range(s, p);
    if (rc != RTB_OK) return rc;
    if (!pixels_rgb) return set_err(RTB_ERR_INVALID, "null pixel buffer");
    CU(cudaSetDevice(s->device));
    const size_t n = (size_t)s->host->cam.width * s->host->cam.height;
    const cudaStream_t stream = cudaStreamPerThread;
    CU(s->accum.reserve(s->device, 4, n * 4 * sizeof(unsigned long long), false));
    CU(cudaMemsetAsync(s->accum.p, 0, n * 4 * sizeof(unsigned long long), stream));
    rc = render_into(s, p, static_cast<unsigned long long*>(s->accum.p), stream);
    if (rc != RTB_OK) return rc;
    rc = accum_to_host(s, static_cast<const unsigned long long*>(s->accum.p), pixels_rgb, stream);
    if (rc != RTB_OK) return rc;
    RtbStats tmp;
    rc = fetch_stats(s, &tmp);
    if (rc != RTB_OK) return rc;
    if (stats) *stats = tmp;
    return RTB_OK;
  });
}

int rtb_render_multi(const RtbSceneDesc* desc, int n_devices, const int* devices, const RtbRenderParams* params, double* pixels_rgb,
                     RtbStats* stats) {
  return guarded([&]() -> int {
    if (!desc || !params || !pixels_rgb) return set_err(RTB_ERR_INVALID, "null argument");
    const int visible = rtb_device_count();
    if (visible == 0) return set_err(RTB_ERR_NO_DEVICE, "no CUDA device visible (this backend has no CPU fallback)");
    if (n_devices < 1 || n_devices > visible) return set_err(RTB_ERR_INVALID, "n_devices outside [1, visible devices]");
    std::vector<int> devs(n_devices);
    for (int k = 0; k < n_devices; k++) devs[k] = devices ? devices[k] : k;
    for (int k = 0; k < n_devices; k++) {
      if (devs[k] < 0 || devs[k] >= visible) return set_err(RTB_ERR_INVALID, "device index out of range");
      for (int j = 0; j < k; j++)
        if (devs[j] == devs[k]) return set_err(RTB_ERR_INVALID, "device listed twice");
    }
    auto host = std::make_shared<HostScene>();
    std::string err;
    int rc = flatten_scene(*desc, *host, err);  // once: every device receives the same arrays
    if (rc != RTB_OK) return set_err(rc, err);
    if (params->sample_begin < 0 || params->sample_end > host->cam.spp || params->sample_begin > params->sample_end)
      return set_err(RTB_ERR_INVALID, "sample range outside [0, spp_used]");
    if (params->pipeline < RTB_PIPELINE_DEFAULT || params->pipeline > RTB_PIPELINE_WAVEFRONT) return set_err(RTB_ERR_INVALID, "unknown pipeline");
    std::vector<Nccl::comm_t> comms;
    if (n_devices > 1) {
      if (!nccl().ok) return set_err(RTB_ERR_UNSUPPORTED, "multi-GPU reduce needs NCCL: " + nccl().why);
      rc = comm_cache().get(devs, comms);
      if (rc != RTB_OK) return rc;
    }
    const size_t n_px = (size_t)host->cam.width * host->cam.height;
    struct Worker {
      rtb_scene* scene = nullptr;
      cudaStream_t stream = nullptr;
      int rc = RTB_OK;
      std::string err;
      RtbStats st{};
    };
    std::vector<Worker> w(n_devices);
    const int64_t s0 = params->sample_begin, total = params->sample_end - params->sample_begin;
    auto work = [&](int k) {  // one host thread per GPU: create, render its contiguous slice of the stratum range
      Worker& me = w[k];
      me.rc = guarded([&]() -> int {
        int r = scene_create(host, devs[k], &me.scene);
        if (r != RTB_OK) return r;
        CU(cudaStreamCreateWithFlags(&me.stream, cudaStreamNonBlocking));
        CU(me.scene->accum.reserve(devs[k], 4, n_px * 4 * sizeof(unsigned long long), false));
        CU(cudaMemsetAsync(me.scene->accum.p, 0, n_px * 4 * sizeof(unsigned long long), me.stream));
        RtbRenderParams p = *params;
        p.sample_begin = s0 + total * k / n_devices;
        p.sample_end = s0 + total * (k + 1) / n_devices;
        r = render_into(me.scene, &p, static_cast<unsigned long long*>(me.scene->accum.p), me.stream);
        if (r != RTB_OK) return r;
        CU(cudaStreamSynchronize(me.stream));
        return fetch_stats(me.scene, &me.st);
      });
      if (me.rc != RTB_OK) me.err = g_err;
    };
    {
      std::vector<std::thread> threads;
      for (int k = 1; k < n_devices; k++) threads.emplace_back(work, k);
      work(0);
      for (std::thread& t : threads) t.join();
    }
    auto cleanup = [&]() {
      for (Worker& me : w) {
        if (me.scene) {
          cudaSetDevice(me.scene->device);
          if (me.stream) { cudaStreamSynchronize(me.stream); cudaStreamDestroy(me.stream); }
          delete me.scene;
        }
        me.scene = nullptr;
      }
    };
    for (const Worker& me : w)
      if (me.rc != RTB_OK) { const int r = me.rc; const std::string m = me.err; cleanup(); return set_err(r, m); }
    // the one collective of the job: int64 sum of the accumulation buffers onto the first device (exact: any
    // split gives the same bits), then ONE device-to-host copy
    if (n_devices > 1) {
      int nrc = nccl().GroupStart();
      for (int k = 0; k < n_devices && nrc == 0; k++) {
        cudaSetDevice(devs[k]);
        nrc = nccl().Reduce(w[k].scene->accum.p, w[k].scene->accum.p, n_px * 4, kNcclInt64, kNcclSum, 0, comms[k], w[k].stream);
      }
      const int erc = nccl().GroupEnd();
      if (nrc == 0) nrc = erc;
      if (nrc != 0) { cleanup(); return set_err(RTB_ERR_CUDA, std::string("ncclReduce: ") + nccl().GetErrorString(nrc)); }
      for (int k = 0; k < n_devices; k++) {
        cudaSetDevice(devs[k]);
        const cudaError_t e = cudaStreamSynchronize(w[k].stream);
        if (e != cudaSuccess) { cleanup(); return cuda_err(e, "reduce"); }
      }
    }
    rc = cudaSetDevice(devs[0]) == cudaSuccess ? RTB_OK : set_err(RTB_ERR_CUDA, "cudaSetDevice");
    if (rc == RTB_OK) rc = accum_to_host(w[0].scene, static_cast<const unsigned long long*>(w[0].scene->accum.p), pixels_rgb, w[0].stream);
    if (rc == RTB_OK && stats) {
      RtbStats t{};
      for (const Worker& me : w) {
        t.paths += me.st.paths; t.segments += me.st.segments; t.node_visits += me.st.node_visits; t.prim_tests += me.st.prim_tests;
        t.medium_probes += me.st.medium_probes; t.nonfinite_samples += me.st.nonfinite_samples; t.kernel_launches += me.st.kernel_launches;
        t.exact_tests += me.st.exact_tests; t.overflow_rays += me.st.overflow_rays;
        t.device_ms = std::max(t.device_ms, me.st.device_ms);
      }
      *stats = t;
    }
    const std::string keep = g_err;
    cleanup();
    if (rc != RTB_OK) g_err = keep;
    return rc;
  });
}

int rtb_camera_rays(const rtb_scene* s, RtbRay* rays) {
  if (!s || !rays) return set_err(RTB_ERR_INVALID, "null argument");
  const DCamera& c = s->host->cam;
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
  return guarded([&]() -> int {
    if (!s || (n > 0 && (!rays || !hits)) || n < 0) return set_err(RTB_ERR_INVALID, "bad argument");
    if (n == 0) return RTB_OK;
    CU(cudaSetDevice(s->device));
    CU(s->scratch_a.reserve((size_t)n * sizeof(RtbRay)));
    CU(s->scratch_b.reserve((size_t)n * sizeof(RtbHit)));
    CU(cudaMemcpy(s->scratch_a.p, rays, (size_t)n * sizeof(RtbRay), cudaMemcpyHostToDevice));
    if (flags & RTB_TRACE_WAVEFRONT) {
      if (flags & RTB_TRACE_BRUTE_FORCE) return set_err(RTB_ERR_INVALID, "RTB_TRACE_WAVEFRONT and RTB_TRACE_BRUTE_FORCE exclude each other");
      if (n > (1ll << 30)) return set_err(RTB_ERR_INVALID, "too many rays for one queue");
      // the f64 rounding bound of the classification scales with the largest coordinate in play: rays of this
      // harness may start anywhere, rays of a render start inside the scene
      double mag = s->dev.scene_mag;
      for (int64_t i = 0; i < n; i++) {
        if (rays[i].t_min != 0.0001) return set_err(RTB_ERR_INVALID, "the wavefront harness traces radiance rays: t_min must be 1e-4");
        mag = std::max(mag, std::fabs(rays[i].origin[0]) + std::fabs(rays[i].origin[1]) + std::fabs(rays[i].origin[2]));
      }
      if (!std::isfinite(mag)) return set_err(RTB_ERR_INVALID, "non-finite ray origin");
      DScene D = s->dev;
      D.scene_mag = std::nextafterf((float)mag, INFINITY);
      const int64_t cap = (n + 255) & ~(int64_t)255;
      CU(s->workspace.reserve(s->device, 2, wavefront_workspace_bytes(D, cap), false));
      CU(s->scratch_c.reserve((size_t)n * sizeof(int)));
      if (const int wrc = wavefront_ready(s)) return wrc;
      unsigned long long overflows = 0;
      CU(launch_trace_wavefront(D, s->wf, s->opt, static_cast<const RtbRay*>(s->scratch_a.p), n, (flags & RTB_TRACE_SECONDARY) != 0,
                                static_cast<RtbHit*>(s->scratch_b.p), static_cast<int*>(s->scratch_c.p), s->workspace.p, s->workspace.bytes,
                                &overflows, cudaStreamPerThread));
      std::vector<int> counts((size_t)n);
      CU(cudaMemcpy(counts.data(), s->scratch_c.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
      s->last = RtbStats{};
      s->last.paths = (uint64_t)n;
      s->last.overflow_rays = overflows;
      for (int c : counts) s->last.exact_tests += (uint64_t)c;
      s->stats_pending = false;
    } else {
      CU(s->scratch_c.reserve(trace_scratch_bytes(n)));
      CU(launch_trace(s->dev, static_cast<const RtbRay*>(s->scratch_a.p), n, flags, static_cast<RtbHit*>(s->scratch_b.p), s->scratch_c.p, 0));
    }
    CU(cudaMemcpy(hits, s->scratch_b.p, (size_t)n * sizeof(RtbHit), cudaMemcpyDeviceToHost));
    return RTB_OK;
  });
}

int rtb_medium_interval(rtb_scene* s, int32_t medium, const RtbRay* rays, int64_t n, double* t_enter, double* t_exit) {
  return guarded([&]() -> int {
    if (!s || n < 0 || (n > 0 && (!rays || !t_enter || !t_exit))) return set_err(RTB_ERR_INVALID, "bad argument");
    if (medium < 0 || medium >= (int)s->host->media.size()) return set_err(RTB_ERR_INVALID, "no such medium");
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
  });
}

int rtb_eval_texture(rtb_scene* s, int32_t texture, const double* uvp, int64_t n, double* rgb_out) {
  return guarded([&]() -> int {
    if (!s || n < 0 || (n > 0 && (!uvp || !rgb_out))) return set_err(RTB_ERR_INVALID, "bad argument");
    if (texture < 0 || texture >= (int)s->host->textures.size()) return set_err(RTB_ERR_INVALID, "no such texture");
    if (n == 0) return RTB_OK;
    CU(cudaSetDevice(s->device));
    CU(s->scratch_a.reserve((size_t)n * 5 * sizeof(double)));
    CU(s->scratch_b.reserve((size_t)n * 3 * sizeof(double)));
    CU(cudaMemcpy(s->scratch_a.p, uvp, (size_t)n * 5 * sizeof(double), cudaMemcpyHostToDevice));
    CU(launch_eval_texture(s->dev, texture, static_cast<const double*>(s->scratch_a.p), n, static_cast<double*>(s->scratch_b.p), 0));
    CU(cudaMemcpy(rgb_out, s->scratch_b.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return RTB_OK;
  });
}

int rtb_eval_light_pdf(rtb_scene* s, const double* origin_dir, int64_t n, double* pdf_out) {
  return guarded([&]() -> int {
    if (!s || n < 0 || (n > 0 && (!origin_dir || !pdf_out))) return set_err(RTB_ERR_INVALID, "bad argument");
    if (s->host->lights.empty()) return set_err(RTB_ERR_INVALID, "scene has no lights");
    if (n == 0) return RTB_OK;
    CU(cudaSetDevice(s->device));
    CU(s->scratch_a.reserve((size_t)n * 6 * sizeof(double)));
    CU(s->scratch_b.reserve((size_t)n * sizeof(double)));
    CU(cudaMemcpy(s->scratch_a.p, origin_dir, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice));
    CU(launch_eval_light_pdf(s->dev, static_cast<const double*>(s->scratch_a.p), n, static_cast<double*>(s->scratch_b.p), 0));
    CU(cudaMemcpy(pdf_out, s->scratch_b.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return RTB_OK;
  });
}

int rtb_write_color(rtb_scene* s, const double* pixels_rgb, int64_t n_pixels, double spp, double exposure, uint8_t* rgb8_out) {
  return guarded([&]() -> int {
    if (n_pixels < 0 || (n_pixels > 0 && (!pixels_rgb || !rgb8_out))) return set_err(RTB_ERR_INVALID, "bad argument");
    if (n_pixels == 0) return RTB_OK;
    if (rtb_device_count() == 0) return set_err(RTB_ERR_NO_DEVICE, "no CUDA device visible (this backend has no CPU fallback)");
    DeviceBuffer tmp_a, tmp_b;  // scene == NULL (the image of a multi-GPU render): the current device, buffers of this call
    DeviceBuffer& a = s ? s->scratch_a : tmp_a;
    DeviceBuffer& b = s ? s->scratch_b : tmp_b;
    if (s) CU(cudaSetDevice(s->device));
    const size_t nv = (size_t)n_pixels * 3;
    CU(a.reserve(nv * sizeof(double)));
    CU(b.reserve(nv));
    CU(cudaMemcpy(a.p, pixels_rgb, nv * sizeof(double), cudaMemcpyHostToDevice));
    CU(launch_write_color(static_cast<const double*>(a.p), (int64_t)nv, spp, exposure, static_cast<uint8_t*>(b.p), 0));
    CU(cudaMemcpy(rgb8_out, b.p, nv, cudaMemcpyDeviceToHost));
    return RTB_OK;
  });
}

// auto_expose  src/render.rs:325-339, literally: medium_weight = 1 / (h * w); medium_point accumulated pixel by pixel;
// divided by spp^2; -ln(0.6) / sqrt(medium_point) above the 0.001 threshold, else 1.
int rtb_auto_expose(const double* pixels_rgb, int64_t n_pixels, double spp, double* exposure_out) {
  if (!pixels_rgb || !exposure_out || n_pixels <= 0 || !(spp > 0.)) return set_err(RTB_ERR_INVALID, "bad argument");
  const double medium_weight = 1. / (double)n_pixels;
  double medium_point = 0.;
  for (int64_t i = 0; i < n_pixels; i++) {
    const double luminance = 0.2126 * pixels_rgb[3 * i] + 0.71516 * pixels_rgb[3 * i + 1] + 0.072169 * pixels_rgb[3 * i + 2];
    medium_point = medium_point + medium_weight * (luminance * luminance);
  }
  medium_point = medium_point / (spp * spp);
  *exposure_out = medium_point > 0.001 ? -std::log(0.6) / std::sqrt(medium_point) : 1.;
  return RTB_OK;
}

int rtb_eval_dielectric(rtb_scene* s, const double* in9, int64_t n, double* dir_out) {
  return guarded([&]() -> int {
    if (!s || n < 0 || (n > 0 && (!in9 || !dir_out))) return set_err(RTB_ERR_INVALID, "bad argument");
    if (n == 0) return RTB_OK;
    CU(cudaSetDevice(s->device));
    CU(s->scratch_a.reserve((size_t)n * 9 * sizeof(double)));
    CU(s->scratch_b.reserve((size_t)n * 3 * sizeof(double)));
    CU(cudaMemcpy(s->scratch_a.p, in9, (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice));
    CU(launch_eval_dielectric(static_cast<const double*>(s->scratch_a.p), n, static_cast<double*>(s->scratch_b.p), 0));
    CU(cudaMemcpy(dir_out, s->scratch_b.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return RTB_OK;
  });
}

int rtb_philox(rtb_scene* s, const uint32_t* ctr_key, int64_t n, uint32_t* out) {
  return guarded([&]() -> int {
    if (!s || n < 0 || (n > 0 && (!ctr_key || !out))) return set_err(RTB_ERR_INVALID, "bad argument");
    if (n == 0) return RTB_OK;
    CU(cudaSetDevice(s->device));
    CU(s->scratch_a.reserve((size_t)n * 6 * sizeof(uint32_t)));
    CU(s->scratch_b.reserve((size_t)n * 4 * sizeof(uint32_t)));
    CU(cudaMemcpy(s->scratch_a.p, ctr_key, (size_t)n * 6 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(launch_philox(static_cast<const uint32_t*>(s->scratch_a.p), n, static_cast<uint32_t*>(s->scratch_b.p), 0));
    CU(cudaMemcpy(out, s->scratch_b.p, (size_t)n * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RTB_OK;
  });
}

}  // extern "C"
