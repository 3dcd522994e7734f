// wavefront.cu -- the wavefront pipeline of the hot path (the product pipeline; the megakernel in
// kernels.cu is the A/B arm).
//
// One iteration advances every in-flight path by one segment:
//
//   k_wf_extend     closest surface hit: persistent warps with dynamic ray     HittableList::hit / BvhNode::hit
//                   fetch and speculative while-while traversal                hittable.rs:88-109, 216-236
//   k_wf_shade      constant-medium events, block-local sort by shading        ConstantMedium::hit  constant_medium.rs:41-95
//                   class, emit / scatter / mixture-pdf sample; survivors      ray_color  render.rs:271-297
//                   are appended (densely) to the next ray queue, finished paths accumulate into the image
//   k_wf_generate   tops the next queue up with new (pixel, stratum) rays      get_ray  render.rs:218-249
//
// Data layout in HBM (DESIGN.md "Queues"): two dense ray queues, each FOUR uint4 PLANES of `capacity`
// entries (SoA of the 64-byte record: origin f64x3 | direction f32x3, time | throughput f32x3, pixel |
// stratum, bounce) in QUEUE ORDER -- there is no slot indirection, every stage streams them and a
// warp's 32 consecutive slots are 512 contiguous bytes per plane; one 16-byte hit record
// {t f64, prim, prim_info.x} per queue position.  All counters live on the device; the host only polls
// "paths left" every few iterations.
//
// Why this shape (ncu, profiles/r01_*): the megakernel keeps only 9.6 of 32 lanes active because
// BVH trip counts differ per ray and lanes sit in different phases; here every kernel is one phase,
// extend refills idle lanes, and shading runs sorted by class.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "rtb_device.cuh"

namespace rtb {

struct WFCounters {
  int n_in;           // rays in the current queue
  int n_out;          // rays appended to the next queue (survivors, then regenerated paths)
  int extend_cursor;  // dynamic-fetch cursor of k_wf_extend
  int n_deferred;     // queue positions the shade kernel left to k_wf_shade_rare this iteration
  unsigned long long next_path;    // camera paths started so far
  unsigned long long total_paths;  // to start in this render call (padded tiles included)
  unsigned long long segments;     // sum of n_in over iterations
  unsigned long long pad1;
};

struct RayRec { uint4 a, b, c, d; };  // the four words of one ray (see pack/unpack); stored as planes, see ray_plane
struct alignas(16) HitRec { double t; int id; int info_x; };  // info_x = prim_info[id].x (0: miss)

struct WFQueues {
  RayRec* rays_a;
  RayRec* rays_b;
  HitRec* hits;
  int* deferred;      // queue positions of deferred items (see k_wf_shade_rare)
  WFCounters* c;
  int capacity;
};

struct PathRec {  // unpacked RayRec
  double ox, oy, oz;
  float dx, dy, dz, time;
  float bx, by, bz;
  uint32_t pixel, sample, bounce;
};

__device__ __forceinline__ RayRec pack(const PathRec& p) {
  RayRec r;
  r.a.x = (unsigned)__double2loint(p.ox); r.a.y = (unsigned)__double2hiint(p.ox);
  r.a.z = (unsigned)__double2loint(p.oy); r.a.w = (unsigned)__double2hiint(p.oy);
  r.b.x = (unsigned)__double2loint(p.oz); r.b.y = (unsigned)__double2hiint(p.oz);
  r.b.z = __float_as_uint(p.dx); r.b.w = __float_as_uint(p.dy);
  r.c.x = __float_as_uint(p.dz); r.c.y = __float_as_uint(p.time);
  r.c.z = __float_as_uint(p.bx); r.c.w = __float_as_uint(p.by);
  r.d.x = __float_as_uint(p.bz); r.d.y = p.pixel; r.d.z = p.sample; r.d.w = p.bounce;
  return r;
}
__device__ __forceinline__ void unpack_geom(const uint4& a, const uint4& b, const uint4& c, PathRec& p) {
  p.ox = __hiloint2double((int)a.y, (int)a.x); p.oy = __hiloint2double((int)a.w, (int)a.z);
  p.oz = __hiloint2double((int)b.y, (int)b.x);
  p.dx = __uint_as_float(b.z); p.dy = __uint_as_float(b.w);
  p.dz = __uint_as_float(c.x); p.time = __uint_as_float(c.y);
  p.bx = __uint_as_float(c.z); p.by = __uint_as_float(c.w);
}
__device__ __forceinline__ void unpack_state(const uint4& d, PathRec& p) {
  p.bz = __uint_as_float(d.x); p.pixel = d.y; p.sample = d.z; p.bounce = d.w;
}
__device__ __forceinline__ Ray to_ray(const PathRec& p) {
  Ray r;
  r.ox = p.ox; r.oy = p.oy; r.oz = p.oz;
  r.dx = (double)p.dx; r.dy = (double)p.dy; r.dz = (double)p.dz; r.time = (double)p.time;
  return r;
}

constexpr uint32_t PADDING_PIXEL = 0xFFFFFFFFu;  // inert lane of a border tile

// 256-bit global load (sm_100: LDG.E.ENL2.256): one instruction per 32-byte sector.  The L1 data pipe
// charges a wavefront per load instruction and distinct sector, and the lanes of a traversing warp sit on
// different nodes -- so a 64-byte node read as 2 x 256 bit costs half the wavefronts of 4 x 128 bit.
struct alignas(32) F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const float4* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p));
  return r;
}

// A ray queue of `cap` slots is stored as FOUR PLANES of cap x 16 bytes (uint4 SoA: plane j holds word j
// of every record), not as cap records of 64 bytes: the 32 consecutive slots a warp reads or writes are
// then 512 contiguous bytes per plane (4 L1 wavefronts per instruction) instead of 32 sectors 64 bytes
// apart (32 wavefronts) -- ncu had 77 % of the shade kernel's L1 wavefronts on these records.
__device__ __forceinline__ const uint4* ray_plane(const RayRec* q, int cap, int j) {
  return reinterpret_cast<const uint4*>(q) + (size_t)j * (size_t)cap;
}
__device__ __forceinline__ uint4* ray_plane(RayRec* q, int cap, int j) {
  return reinterpret_cast<uint4*>(q) + (size_t)j * (size_t)cap;
}

// Queue records are streamed exactly once per kernel: load/store them with the evict-first policy so
// that they do not push the BVH nodes and primitives (re-read by every ray) out of L1/L2.
#if defined(RTB_NO_STREAM_HINTS)
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldg(p); }
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) { *p = v; }
#else
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) { __stcs(p, v); }
#endif

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t queue_bytes(int64_t n) {
  return 2 * align_up(n * sizeof(RayRec)) + align_up(n * sizeof(HitRec)) + align_up(n * sizeof(int)) + align_up(sizeof(WFCounters));
}
// traversal stacks of the pool extend kernel: one int[WF_POOL_STACK] per pool entry of every resident warp
constexpr int WF_POOL_MAX_BLOCKS_PER_SM = 12;
static size_t pool_scratch_bytes(int sms) { return align_up((size_t)sms * WF_POOL_MAX_BLOCKS_PER_SM * (128 / 32) * 64 * 32 * sizeof(int)); }
// room for up to WF_MAX_SUB sub-pipelines that share `n` path slots (alignment slack per sub-pipeline)
size_t wavefront_workspace_bytes(const DScene&, int64_t n) { return queue_bytes(n) + WF_MAX_SUB * 4096 + pool_scratch_bytes(160); }

static WFQueues carve(void* ws, int64_t n) {
  char* p = static_cast<char*>(ws);
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
  WFQueues q;
  q.rays_a = reinterpret_cast<RayRec*>(take(n * sizeof(RayRec)));
  q.rays_b = reinterpret_cast<RayRec*>(take(n * sizeof(RayRec)));
  q.hits = reinterpret_cast<HitRec*>(take(n * sizeof(HitRec)));
  q.deferred = reinterpret_cast<int*>(take(n * sizeof(int)));
  q.c = reinterpret_cast<WFCounters*>(take(sizeof(WFCounters)));
  q.capacity = (int)n;
  return q;
}

// ------------------------------------------------------------------------------------------------
// bookkeeping (one thread)
// ------------------------------------------------------------------------------------------------
__global__ void k_wf_init(WFQueues Q, unsigned long long total_paths) {
  WFCounters z = {};
  z.total_paths = total_paths;
  *Q.c = z;
}

// after shade + generate: the out queue becomes the in queue of the next iteration
__global__ void k_wf_advance(WFQueues Q) {
  WFCounters* c = Q.c;
  const unsigned long long left = c->total_paths - c->next_path;
  const unsigned long long room = (unsigned long long)(Q.capacity - c->n_out);
  const unsigned long long gen = left < room ? left : room;
  c->next_path += gen;
  c->n_in = c->n_out + (int)gen;
  c->segments += (unsigned long long)c->n_in;
  c->n_out = 0;
  c->extend_cursor = 0;
  c->n_deferred = 0;
}

// ------------------------------------------------------------------------------------------------
// generate: path id -> (stratum, pixel).  Ids enumerate 8x4 pixel tiles padded to 32 lanes, so the
// 32 consecutive rays of one warp-fetch in extend are one coherent tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_generate(const __grid_constant__ DScene S, WFQueues Q, long long s_begin,
                                                      RayRec* __restrict__ out) {
  const WFCounters* c = Q.c;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long left = c->total_paths - c->next_path;
  const int n_out = c->n_out;
  if (i >= Q.capacity - n_out || (unsigned long long)i >= left) return;
  const unsigned long long pid = c->next_path + (unsigned long long)i;
  const uint32_t tiles_x = (uint32_t)(S.cam.width + 7) >> 3;
  const uint32_t tiles = tiles_x * (uint32_t)((S.cam.height + 3) >> 2);
  const unsigned long long padded = (unsigned long long)tiles * 32ull;
  const uint32_t sample = (uint32_t)(s_begin + (long long)(pid / padded));
  const uint32_t idx = (uint32_t)(pid % padded), tile = idx >> 5, lane = idx & 31u;
  const uint32_t x = (tile % tiles_x) * 8u + (lane & 7u), y = (tile / tiles_x) * 4u + (lane >> 3);
  PathRec p;
  if (x < (uint32_t)S.cam.width && y < (uint32_t)S.cam.height) {
    const uint32_t pixel = y * (uint32_t)S.cam.width + x;
    PathState ps;
    generate_primary(S, pixel, sample, ps);
    p.ox = ps.ray.ox; p.oy = ps.ray.oy; p.oz = ps.ray.oz;
    p.dx = (float)ps.ray.dx; p.dy = (float)ps.ray.dy; p.dz = (float)ps.ray.dz; p.time = (float)ps.ray.time;
    p.bx = p.by = p.bz = 1.f;
    p.pixel = pixel;
  } else {  // padding lane: dies in its first shade without touching the image
    p.ox = p.oy = p.oz = 0.;
    p.dx = p.dy = 0.f; p.dz = 1.f; p.time = 0.f;
    p.bx = p.by = p.bz = 0.f;
    p.pixel = PADDING_PIXEL;
  }
  p.sample = sample;
  p.bounce = 0u;
  const RayRec rec = pack(p);
  const int o = n_out + i, cap = Q.capacity;
  st_stream(ray_plane(out, cap, 0) + o, rec.a); st_stream(ray_plane(out, cap, 1) + o, rec.b);
  st_stream(ray_plane(out, cap, 2) + o, rec.c); st_stream(ray_plane(out, cap, 3) + o, rec.d);
}

// ------------------------------------------------------------------------------------------------
// extend: persistent warps, dynamic fetch, speculative while-while traversal (Aila & Laine 2009):
// a lane that reaches a leaf postpones it and keeps descending until every lane of the warp holds a
// leaf or has run out of nodes; then the leaves are tested together.  fp32 conservative slabs in
// fused form (t = plane * inv_d - o * inv_d), f64 reference-order primitive tests.
// ------------------------------------------------------------------------------------------------
#ifndef WF_EXTEND_BLOCK_DIM
#define WF_EXTEND_BLOCK_DIM 128
#endif
constexpr int WF_EXTEND_BLOCK = WF_EXTEND_BLOCK_DIM;
#ifndef WF_EXTEND_MIN_BLOCKS
#define WF_EXTEND_MIN_BLOCKS (1024 / WF_EXTEND_BLOCK_DIM)  // 32 warps per SM = 64 registers, no spills (measured: 28 warps 28.7, 32 warps 27.3 ms per c4 row)
#endif
#ifndef WF_FETCH_THRESHOLD_N
#define WF_FETCH_THRESHOLD_N 20  // measured on c4: 32 -> 26.9, 28 -> 26.8, 24 -> 26.7, 20 -> 26.5 ms extend per row
#endif
constexpr int WF_FETCH_THRESHOLD = WF_FETCH_THRESHOLD_N;  // refill when fewer than this many lanes hold a ray
constexpr int TRAV_DONE = 0x7FFFFFFF;
#ifndef WF_LD256
#define WF_LD256 1
#endif
#ifndef WF_BREAK_LEFT
#define WF_BREAK_LEFT 16  // measured on c4: 0 -> 35.1, 12 -> 33.9, 16 -> 33.4, 20 -> 33.5, 24 -> 33.8 ms extend per step
#endif

// NODES: which form of the tree is traversed, chosen per scene by the builder --
//   NODES_BVH2  the 64-byte fp32 nodes;
//   NODES_Q     the 32-byte quantised nodes (rtb_device.cuh, slab_box_q; opt-in, DScene::use_qnodes);
//   NODES_BVH4  every other level collapsed (DScene::use_bvh4): half the dependent steps per ray.
//   NODES_BVH2_MULTI  the BVH2 with leaves of several primitives (RTB_BVH_LEAF > 1): the generic leaf loop compiled in.
enum : int { NODES_BVH2 = 0, NODES_Q = 1, NODES_BVH4 = 2, NODES_BVH2_MULTI = 3 };
template <bool STATS, int NODES>
__global__ void __launch_bounds__(WF_EXTEND_BLOCK, WF_EXTEND_MIN_BLOCKS) k_wf_extend(const __grid_constant__ DScene S, WFQueues Q,
                                                               const RayRec* __restrict__ rays_in,
                                                               DStats* __restrict__ stats) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int n = Q.c->n_in;
  unsigned long long st_nodes = 0, st_prims = 0;
  bool have = false, exhausted = false;
  int pos = -1;
  // the ray waits for the f64 leaf tests as stored (f64 origin, fp32 direction and time): fewer live registers
  double rox = 0., roy = 0., roz = 0.;
  float rdx = 0.f, rdy = 0.f, rdz = 0.f, rtime = 0.f;
  SlabRay sr = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float tbest32 = 0.f;
  const float tmin32 = __double2float_rd(0.0001);
  Hit best;
  hit_reset(best);
  int stack[BVH_STACK];
  // (A stale-entry cull -- stacking each subtree's entry distance and dropping entries beyond the best
  // hit unvisited -- was measured: 0.4 fewer visits per ray, but the doubled stack traffic made it slower.)
#define WF_POP() node = sp > 0 ? stack[--sp] : TRAV_DONE
  int sp = 0, node = TRAV_DONE, leaf = 0;  // leaf: postponed leaf reference (< 0) or 0
  const bool any_surface = S.n_surface_prims > 0;
  for (;;) {
    // ---- dynamic fetch -------------------------------------------------------------------------
    const unsigned have_mask = __ballot_sync(FULL, have);
    if (__popc(have_mask) < WF_FETCH_THRESHOLD) {
      const unsigned need = __ballot_sync(FULL, !have && !exhausted);
      if (need) {
        int base = 0;
        const int leader = __ffs(need) - 1;
        if (lane == leader) base = atomicAdd(&Q.c->extend_cursor, __popc(need));
        base = __shfl_sync(FULL, base, leader);
        if (!have && !exhausted) {
          const int k = base + __popc(need & ((1u << lane) - 1u));
          if (k < n) {
            pos = k;
            const uint4 a = ld_stream(ray_plane(rays_in, Q.capacity, 0) + k), b = ld_stream(ray_plane(rays_in, Q.capacity, 1) + k),
                        c = ld_stream(ray_plane(rays_in, Q.capacity, 2) + k);
            PathRec p;
            unpack_geom(a, b, c, p);
            rox = p.ox; roy = p.oy; roz = p.oz;
            rdx = p.dx; rdy = p.dy; rdz = p.dz; rtime = p.time;
            sr = NODES == NODES_Q ? slab_ray_q(S, p.ox, p.oy, p.oz, p.dx, p.dy, p.dz) : slab_ray(p.ox, p.oy, p.oz, p.dx, p.dy, p.dz);
            hit_reset(best);
            tbest32 = __double2float_ru(best.t);
            sp = 0;
            leaf = 0;
            node = any_surface ? 0 : TRAV_DONE;
            have = true;
          } else {
            exhausted = true;
          }
        }
      }
    }
    if (!__any_sync(FULL, have)) break;
    // ---- inner nodes, speculative: keep descending after the first leaf is found -------------------
#if WF_BREAK_LEFT > 0
    const int entered = __popc(__ballot_sync(FULL, node >= 0 && node != TRAV_DONE));
#endif
    while (node >= 0 && node != TRAV_DONE) {
      if (STATS) st_nodes++;
      if (NODES == NODES_BVH4) {
        const float4* N = S.nodes4 + 8 * (size_t)node;
#if WF_LD256
        const F8 A = ldg256(N + 0), B = ldg256(N + 2), C = ldg256(N + 4), D = ldg256(N + 6);
        const float4 lx = A.lo, hx = A.hi, ly = B.lo, hy = B.hi, lz = C.lo, hz = C.hi;
        const int4 rf = make_int4(__float_as_int(D.lo.x), __float_as_int(D.lo.y), __float_as_int(D.lo.z), __float_as_int(D.lo.w));
#else
        const float4 lx = __ldg(N + 0), hx = __ldg(N + 1), ly = __ldg(N + 2), hy = __ldg(N + 3), lz = __ldg(N + 4), hz = __ldg(N + 5);
        const int4 rf = __ldg(reinterpret_cast<const int4*>(N + 6));
#endif
        float t0, t1, t2, t3;
        bool h0, h1, h2, h3;
        slab_box(lx.x, hx.x, ly.x, hy.x, lz.x, hz.x, sr, tmin32, tbest32, t0, h0);
        slab_box(lx.y, hx.y, ly.y, hy.y, lz.y, hz.y, sr, tmin32, tbest32, t1, h1);
        slab_box(lx.z, hx.z, ly.z, hy.z, lz.z, hz.z, sr, tmin32, tbest32, t2, h2);
        slab_box(lx.w, hx.w, ly.w, hy.w, lz.w, hz.w, sr, tmin32, tbest32, t3, h3);
        // nearest child next, the other hits are stacked in slot order (the warp model shows a full
        // near-to-far sort buys < 1 % fewer visits)
        const float inf = __int_as_float(0x7F800000);
        const float k0 = h0 ? t0 : inf, k1 = h1 ? t1 : inf, k2 = h2 ? t2 : inf, k3 = h3 ? t3 : inf;
        const float kmin = fminf(fminf(k0, k1), fminf(k2, k3));
        if (h0 | h1 | h2 | h3) {
          const bool n0 = h0 && k0 == kmin, n1 = !n0 && h1 && k1 == kmin, n2 = !(n0 | n1) && h2 && k2 == kmin;
          const bool n3 = !(n0 | n1 | n2);
          if (h0 && !n0) stack[sp++] = rf.x;
          if (h1 && !n1) stack[sp++] = rf.y;
          if (h2 && !n2) stack[sp++] = rf.z;
          if (h3 && !n3) stack[sp++] = rf.w;
          node = n0 ? rf.x : (n1 ? rf.y : (n2 ? rf.z : rf.w));
        } else {
          WF_POP();
        }
      } else {
      float tn0, tn1;
      bool h0, h1;
      int ch0, ch1;
      if (NODES == NODES_Q) {
        const uint4* N = S.qnodes + 2 * (size_t)node;
        const uint4 q0 = __ldg(N + 0), q1 = __ldg(N + 1);
        slab_box_q(q0.x, q0.y, q0.z, sr, tmin32, tbest32, tn0, h0);
        slab_box_q(q1.x, q1.y, q1.z, sr, tmin32, tbest32, tn1, h1);
        ch0 = (int)q0.w; ch1 = (int)q1.w;
      } else {
        const float4* N = S.nodes + 4 * (size_t)node;
#if WF_LD256
        const F8 A = ldg256(N + 0), B = ldg256(N + 2);
        const float4 n0 = A.lo, n1 = A.hi, n2 = B.lo, n3 = B.hi;
#else
        const float4 n0 = __ldg(N + 0), n1 = __ldg(N + 1), n2 = __ldg(N + 2), n3 = __ldg(N + 3);
#endif
        slab_box(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, sr, tmin32, tbest32, tn0, h0);
        slab_box(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, sr, tmin32, tbest32, tn1, h1);
        ch0 = __float_as_int(n3.x); ch1 = __float_as_int(n3.y);
      }
      if (h0 && h1) {
        if (tn1 < tn0) { const int tmp = ch0; ch0 = ch1; ch1 = tmp; }
        stack[sp++] = ch1;
        node = ch0;
      } else if (h0) {
        node = ch0;
      } else if (h1) {
        node = ch1;
      } else {
        WF_POP();
      }
      }
      if (node < 0 && leaf == 0) {  // first leaf: postpone it and continue with the next node
        leaf = node;  // (prefetching its primitive here was measured: slower -- the L1 data pipe is the scarce resource)
        WF_POP();
      }
      // every lane still in this loop holds a leaf already: stop speculating and test the leaves.
      // WF_BREAK_LEFT: also stop once that many lanes have dropped out of the loop (second leaf, or
      // out of nodes) -- the warp model (tools/sim) puts the inner loop at 25 instead of 18 lanes.
      const unsigned looping = __activemask();
      if (!__any_sync(looping, leaf == 0)) break;
#if WF_BREAK_LEFT > 0
      if (entered - __popc(looping) >= WF_BREAK_LEFT) break;
#endif
    }
    // ---- leaves: the postponed one, then the current node if it is a leaf too -------------------------
    Ray r;
    r.ox = rox; r.oy = roy; r.oz = roz;
    r.dx = (double)rdx; r.dy = (double)rdy; r.dz = (double)rdz; r.time = (double)rtime;
    while (leaf < 0) {
      const int count = test_leaf<NODES == NODES_BVH2_MULTI>(S, leaf, r, 0.0001, best);
      if (STATS) st_prims += (unsigned long long)count;
      leaf = 0;
      if (node < 0) {
        leaf = node;
        WF_POP();
      }
    }
    if (have) tbest32 = __double2float_ru(best.t);
    if (have && node == TRAV_DONE) {
      uint4 h;  // {t, prim, prim_info.x}: kind | flags | class | material travel with the hit
      h.x = (unsigned)__double2loint(best.t); h.y = (unsigned)__double2hiint(best.t);
      h.z = (unsigned)best.prim; h.w = best.prim >= 0 ? (unsigned)__ldg(&S.prim_info[best.prim].x) : 0u;
      st_stream(reinterpret_cast<uint4*>(Q.hits + pos), h);
      have = false;
    }
  }
  if (STATS) {
    atomicAdd(&stats->node_visits, st_nodes);
    atomicAdd(&stats->prim_tests, st_prims);
  }
}

// ------------------------------------------------------------------------------------------------
// extend, pool variant.  ncu on k_wf_extend: 15 of 32 lanes active -- the inner loop waits for the
// slowest descent of the warp (18 lanes) and the f64 leaf tests run split by primitive kind (6-9
// lanes).  Here each persistent warp keeps WF_POOL rays (2 per lane) with their traversal state in
// shared memory; every iteration the warp picks ONE phase -- inner-node steps, sphere leaves or quad
// leaves, whichever has most entries ready -- and its 32 lanes take up to 32 entries in that phase.
// Lanes are therefore nearly always full and a leaf pass tests one primitive kind only.  Traversal
// stacks live in a per-warp global scratch (L1-resident, like local memory).
// Results are bit-identical to k_wf_extend: the same tests, and closest hits do not depend on order.
// MEASURED (profiles/r01_pool_extend.txt): lanes rise 15 -> 19 (leaf tests 6 -> 21) and warp
// instructions drop 10 %, but the pick bookkeeping (ballots, shared-memory state round trips) and the
// loss of back-to-back node loads cost more: 42.9 ms vs 34.5 ms per c4 step.  Kept as an opt-in
// (RTB_WF_EXTEND_POOL=1) and as a parity cross-check of the default kernel.
// ------------------------------------------------------------------------------------------------
constexpr int WF_POOL = 64;        // entries per warp
constexpr int WF_POOL_STACK = 32;  // stack entries per ray (deeper trees fall back to k_wf_extend)
constexpr int WF_POOL_BLOCK = 128;
constexpr int WF_POOL_WARPS = WF_POOL_BLOCK / 32;
#ifndef WF_POOL_CHUNK_STEPS
#define WF_POOL_CHUNK_STEPS 6
#endif
constexpr int WF_POOL_CHUNK = WF_POOL_CHUNK_STEPS;   // inner-node steps per pick
enum : int { ST_EMPTY = 0, ST_INNER = 1, ST_LEAF_S = 2, ST_LEAF_Q = 3 };

struct PoolWarp {
  double ox[WF_POOL], oy[WF_POOL], oz[WF_POOL], tbest[WF_POOL];
  float dx[WF_POOL], dy[WF_POOL], dz[WF_POOL], time[WF_POOL];
  float idx[WF_POOL], idy[WF_POOL], idz[WF_POOL], oxi[WF_POOL], oyi[WF_POOL], ozi[WF_POOL];
  int best_prim[WF_POOL];
  int node[WF_POOL], sp[WF_POOL], pos[WF_POOL], status[WF_POOL];
  int list[32];
};

__device__ __forceinline__ int leaf_status(const DScene& S, int node) {
  // kind of the (first) primitive of a leaf; mixed leaves are tested by test_prim anyway
  return (leaf_kind_bits(node) & LEAF_KIND_QUAD) ? ST_LEAF_Q : ST_LEAF_S;
}

template <bool STATS>
__global__ void __launch_bounds__(WF_POOL_BLOCK) k_wf_extend_pool(const __grid_constant__ DScene S, WFQueues Q,
                                                                 const RayRec* __restrict__ rays_in,
                                                                 int* __restrict__ stack_scratch,
                                                                 DStats* __restrict__ stats) {
  const unsigned FULL = 0xFFFFFFFFu;
  __shared__ PoolWarp pool[WF_POOL_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PoolWarp& W = pool[warp];
  int* const stacks = stack_scratch + ((size_t)(blockIdx.x * WF_POOL_WARPS + warp) * WF_POOL) * WF_POOL_STACK;
  const int n = Q.c->n_in;
  const unsigned lt = (1u << lane) - 1u;
  const float tmin32 = __double2float_rd(0.0001);
  const bool any_surface = S.n_surface_prims > 0;
  unsigned long long st_nodes = 0, st_prims = 0;
  W.status[lane] = ST_EMPTY;
  W.status[lane + 32] = ST_EMPTY;
  bool exhausted = false;
  __syncwarp();
  for (;;) {
    int s0 = W.status[lane], s1 = W.status[lane + 32];
    // ---- refill empty entries from the queue ---------------------------------------------------------
    {
      const unsigned e0 = __ballot_sync(FULL, s0 == ST_EMPTY), e1 = __ballot_sync(FULL, s1 == ST_EMPTY);
      const int n_empty = __popc(e0) + __popc(e1);
      if (!exhausted && n_empty >= 16) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&Q.c->extend_cursor, n_empty);
        base = __shfl_sync(FULL, base, 0);
        exhausted = base + n_empty >= n;
#pragma unroll
        for (int half = 0; half < 2; half++) {
          const bool mine = half == 0 ? (s0 == ST_EMPTY) : (s1 == ST_EMPTY);
          const int k = base + (half == 0 ? __popc(e0 & lt) : __popc(e0) + __popc(e1 & lt));
          if (mine && k < n) {
            const int e = lane + 32 * half;
            const uint4 a = __ldg(ray_plane(rays_in, Q.capacity, 0) + k), b = __ldg(ray_plane(rays_in, Q.capacity, 1) + k),
                        c = __ldg(ray_plane(rays_in, Q.capacity, 2) + k);
            PathRec p;
            unpack_geom(a, b, c, p);
            if (any_surface) {
              const float ix = safe_rcp(p.dx), iy = safe_rcp(p.dy), iz = safe_rcp(p.dz);
              W.ox[e] = p.ox; W.oy[e] = p.oy; W.oz[e] = p.oz;
              W.dx[e] = p.dx; W.dy[e] = p.dy; W.dz[e] = p.dz; W.time[e] = p.time;
              W.idx[e] = ix; W.idy[e] = iy; W.idz[e] = iz;
              W.oxi[e] = (float)p.ox * ix; W.oyi[e] = (float)p.oy * iy; W.ozi[e] = (float)p.oz * iz;
              W.tbest[e] = RTB_INF;
              W.best_prim[e] = -1;
              W.node[e] = 0; W.sp[e] = 0; W.pos[e] = k;
              W.status[e] = ST_INNER;
            } else {  // nothing to traverse: record the miss
              HitRec h;
              h.t = RTB_INF; h.id = -1; h.info_x = 0;
              Q.hits[k] = h;
            }
          }
        }
        __syncwarp();
        s0 = W.status[lane]; s1 = W.status[lane + 32];
      }
    }
    // ---- pick the phase with most ready entries ---------------------------------------------------------
    const unsigned i0 = __ballot_sync(FULL, s0 == ST_INNER), i1 = __ballot_sync(FULL, s1 == ST_INNER);
    const unsigned p0 = __ballot_sync(FULL, s0 == ST_LEAF_S), p1 = __ballot_sync(FULL, s1 == ST_LEAF_S);
    const unsigned q0 = __ballot_sync(FULL, s0 == ST_LEAF_Q), q1 = __ballot_sync(FULL, s1 == ST_LEAF_Q);
    const int nI = __popc(i0) + __popc(i1), nS = __popc(p0) + __popc(p1), nQ = __popc(q0) + __popc(q1);
    if (nI + nS + nQ == 0) {
      if (exhausted) break;
      continue;  // everything empty but rays remain: the refill above runs next turn (n_empty == 64)
    }
    int phase;
    unsigned m0, m1;
    if (nI >= 32 || (nI >= nS && nI >= nQ)) { phase = ST_INNER; m0 = i0; m1 = i1; }
    else if (nS >= nQ) { phase = ST_LEAF_S; m0 = p0; m1 = p1; }
    else { phase = ST_LEAF_Q; m0 = q0; m1 = q1; }
    {
      const int r0 = __popc(m0 & lt), r1 = __popc(m0) + __popc(m1 & lt);
      if (((m0 >> lane) & 1u) && r0 < 32) W.list[r0] = lane;
      if (((m1 >> lane) & 1u) && r1 < 32) W.list[r1] = lane + 32;
    }
    const int cnt = min(32, __popc(m0) + __popc(m1));
    __syncwarp();
    if (lane < cnt) {
      const int e = W.list[lane];
      int* const stack = stacks + e * WF_POOL_STACK;
      int node = W.node[e], sp = W.sp[e];
      if (phase == ST_INNER) {
        const SlabRay sr = {W.idx[e], W.idy[e], W.idz[e], W.oxi[e], W.oyi[e], W.ozi[e]};
        const float tbest32 = __double2float_ru(W.tbest[e]);
#pragma unroll 1
        for (int step = 0; step < WF_POOL_CHUNK && node >= 0 && node != TRAV_DONE; step++) {
          if (STATS) st_nodes++;
          const float4* N = S.nodes + 4 * (size_t)node;
          const float4 n0 = __ldg(N + 0), n1 = __ldg(N + 1), n2 = __ldg(N + 2), n3 = __ldg(N + 3);
          float tn0, tn1;
          bool h0, h1;
          slab_box(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, sr, tmin32, tbest32, tn0, h0);
          slab_box(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, sr, tmin32, tbest32, tn1, h1);
          int ch0 = __float_as_int(n3.x), ch1 = __float_as_int(n3.y);
          if (h0 && h1) {
            if (tn1 < tn0) { const int tmp = ch0; ch0 = ch1; ch1 = tmp; }
            stack[sp++] = ch1;
            node = ch0;
          } else if (h0) {
            node = ch0;
          } else if (h1) {
            node = ch1;
          } else {
            node = sp > 0 ? stack[--sp] : TRAV_DONE;
          }
        }
      } else {
        // ---- leaf: f64 reference-order tests of one primitive kind ----------------------------------------
        Ray r;
        r.ox = W.ox[e]; r.oy = W.oy[e]; r.oz = W.oz[e];
        r.dx = (double)W.dx[e]; r.dy = (double)W.dy[e]; r.dz = (double)W.dz[e]; r.time = (double)W.time[e];
        Hit best;
        best.t = W.tbest[e]; best.a = 0.; best.b = 0.;
        best.prim = W.best_prim[e];
        const int count = test_leaf(S, node, r, 0.0001, best);
        if (STATS) st_prims += (unsigned long long)count;
        W.tbest[e] = best.t;
        W.best_prim[e] = best.prim;
        node = sp > 0 ? stack[--sp] : TRAV_DONE;
      }
      // ---- write the entry back / retire it -------------------------------------------------------------------
      if (node == TRAV_DONE) {
        HitRec h;
        h.t = W.tbest[e]; h.id = W.best_prim[e];
        h.info_x = h.id >= 0 ? __ldg(&S.prim_info[h.id].x) : 0;
        Q.hits[W.pos[e]] = h;
        W.status[e] = ST_EMPTY;
      } else {
        W.node[e] = node;
        W.sp[e] = sp;
        W.status[e] = node >= 0 ? ST_INNER : leaf_status(S, node);
      }
    }
    __syncwarp();
  }
  if (STATS) {
    atomicAdd(&stats->node_visits, st_nodes);
    atomicAdd(&stats->prim_tests, st_prims);
  }
}

// ------------------------------------------------------------------------------------------------
// shade (fused resolve + sort + shade).  One block owns WF_SHADE_BLOCK consecutive queue positions:
//   1. stream the ray + hit records in (coalesced), add the constant-medium events
//      (ConstantMedium::hit, constant_medium.rs:41-95) and look up the shading class;
//   2. block-local counting sort by class through shared memory, so that a warp shades one class
//      (the Perlin texture costs ~1k instructions, a solid Lambertian ~150: they must not share a warp);
//   3. shade (ray_color's match arms, render.rs:271-297); survivors are appended densely to the next
//      queue with one atomic per block, finished paths add their radiance to the image.
// No global bins, no gathers: every global access of this kernel is sequential.
// ------------------------------------------------------------------------------------------------
#ifndef WF_SHADE_BLOCK_DIM
#define WF_SHADE_BLOCK_DIM 128  // measured on c4: 128 -> 23.0, 256 -> 23.9, 512 -> 26.0 ms shade per step
#endif
#ifndef WF_SHADE_MIN_BLOCKS
#define WF_SHADE_MIN_BLOCKS 7  // 72 registers (148 B of spills): measured 20.3 vs 21.1 ms per c4 row at 6 blocks / 80 registers
#endif
constexpr int WF_SHADE_BLOCK = WF_SHADE_BLOCK_DIM;

struct ShadeItem {  // what moves through shared memory to the lane that shades it (80 B)
  uint4 a, b, c, d;
  double t;
  int id;
  int info_x;  // prim_info[id].x of a surface hit (kind | flags | class | material)
};
// ---- the three per-item pieces shared by both shade kernels --------------------------------------------------
// 1. constant-medium events (ConstantMedium::hit, constant_medium.rs:41-95), each lane for its own ray, then
//    the shading class.  The f64 boundary intervals run at ~13 of 32 lanes (only the rays whose fp32
//    rejections do not settle it); queueing those (ray, medium) pairs in shared memory and evaluating the
//    queue densely after a barrier was measured: 18.8 -> 23.1 ms per c4 row (a 4-warp block idles through
//    the whole f64 chain).  Kept per-lane.
template <int SPEC>
__device__ __forceinline__ int wf_resolve(const DScene& S, const Tables& T, const uint4& a, const uint4& b, const uint4& c,
                                          const uint4& d, double& t, int& id, int info_x) {
  if (d.y == PADDING_PIXEL) return CLS_MISS;
  if ((SPEC & SPEC_MEDIA) && S.n_media > 0) {
    PathRec p;
    unpack_geom(a, b, c, p);
    unpack_state(d, p);
    Rand4 u;
    for (int mi = 0; mi < S.n_media; mi++) {
      if ((mi & 3) == 0) u = rand4(S, p.pixel, p.sample, p.bounce, 1u + (uint32_t)(mi >> 2));
      const float U = (mi & 3) == 0 ? u.x : ((mi & 3) == 1 ? u.y : ((mi & 3) == 2 ? u.z : u.w));
      const double tm = medium_event_lazy<(SPEC & SPEC_BOXSCAN) != 0, (SPEC & SPEC_GENERIC_MEDIA) != 0>(S, T.media[mi], p.ox, p.oy, p.oz, p.dx, p.dy, p.dz, (double)p.time, 0.0001, t, U);
      if (tm < t) { t = tm; id = -2 - mi; }
    }
  }
  return id == -1 ? CLS_MISS : (id >= 0 ? ((info_x >> PRIM_CLASS_SHIFT) & 0xF) : (T.media[-2 - id].cls_fast & 0xF));
}

// 2. the block-local counting sort by class: position of this lane's item among the block's items, before the
//    class prefix is added (one shared-memory atomic per warp and class present).
//    (measured alternatives: per-class ballots + a prefix pass, 23.9 vs 25.1 ms per c4 row; no sort at all
//     25.2 ms and a less coherent next queue -- extend 34.6 vs 33.4 ms; scene tables staged in shared memory +1.8 ms)
__device__ __forceinline__ int wf_class_slot(int cls, int lane, int* class_count) {
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, cls);
  int warp_base = 0;
  const int leader = __ffs(peers) - 1;
  if (cls >= 0 && lane == leader) warp_base = atomicAdd(&class_count[cls], __popc(peers));
  warp_base = __shfl_sync(0xFFFFFFFFu, warp_base, leader);
  return warp_base + __popc(peers & ((1u << lane) - 1u));
}
__device__ __forceinline__ int wf_class_prefix(int cls, const int* class_count) {
  int base = 0;
#pragma unroll
  for (int k = 0; k < NUM_CLASSES - 1; k++)
    if (k < cls) base += class_count[k];
  return base;
}

// 3. ray_color's match arms for one item (render.rs:271-297): finished paths add their radiance to the image,
//    survivors return true with the next ray packed into `out`.
template <bool STATS, int SPEC>
__device__ __forceinline__ bool wf_shade_item(const DScene& S, const Tables& T, const uint4& a, const uint4& b, const uint4& c,
                                              const uint4& d, double t, int id, int info_x, float4* __restrict__ accum,
                                              RayRec& out, DStats& st) {
  if (d.y == PADDING_PIXEL) return false;
  PathRec p;
  unpack_geom(a, b, c, p);
  unpack_state(d, p);
  PathState ps;
  ps.ray = to_ray(p);
  ps.bx = p.bx; ps.by = p.by; ps.bz = p.bz;
  ps.pixel = p.pixel; ps.sample = p.sample; ps.bounce = p.bounce;
  Event ev;
  ev.t = t; ev.a = 0.; ev.b = 0.; ev.have_ab = 0;
  ev.prim = id >= 0 ? id : -1;
  ev.medium = id <= -2 ? -2 - id : -1;
  ev.info_x = id >= 0 ? info_x : 0;
  float Lr = 0.f, Lg = 0.f, Lb = 0.f;
  if (shade<(SPEC & SPEC_LIGHTS) != 0, (SPEC & SPEC_QUAD_UV) != 0, (SPEC & SPEC_SPHERE_UV) != 0, (SPEC & SPEC_TEXTURES) != 0>(S, T, ps, ev, Lr, Lg,
                                                                                                                          Lb, &st, STATS)) {
    p.ox = ps.ray.ox; p.oy = ps.ray.oy; p.oz = ps.ray.oz;
    p.dx = (float)ps.ray.dx; p.dy = (float)ps.ray.dy; p.dz = (float)ps.ray.dz;
    p.bx = ps.bx; p.by = ps.by; p.bz = ps.bz;
    p.bounce = ps.bounce;
    out = pack(p);
    return true;
  }
  const bool finite = (fabsf(Lr) < 3.0e38f) && (fabsf(Lg) < 3.0e38f) && (fabsf(Lb) < 3.0e38f);
  float* acc = reinterpret_cast<float*>(accum + p.pixel);
  if (finite || (S.flags & 2u)) {
    if (Lr != 0.f) atomicAdd(acc + 0, Lr);
    if (Lg != 0.f) atomicAdd(acc + 1, Lg);
    if (Lb != 0.f) atomicAdd(acc + 2, Lb);
  } else if (STATS) {
    st.nonfinite++;
  }
  // (the per-pixel sample count, accum.w, is added in bulk by k_wf_add_count: every pixel receives
  //  exactly one path per stratum, so one atomic per path would only repeat what the host knows)
  return false;
}

// survivors: one atomic per warp, dense coalesced append.  (A per-block aggregate would need a barrier
// after shading, and ncu showed every warp then waits for the block's slowest class.)
__device__ __forceinline__ void wf_append(const WFQueues& Q, RayRec* __restrict__ rays_out, bool alive, const RayRec& out, int lane) {
  const unsigned m = __ballot_sync(0xFFFFFFFFu, alive);
  if (!m) return;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&Q.c->n_out, __popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (alive) {
    const int o = base + __popc(m & ((1u << lane) - 1u)), cap = Q.capacity;
    st_stream(ray_plane(rays_out, cap, 0) + o, out.a); st_stream(ray_plane(rays_out, cap, 1) + o, out.b);
    st_stream(ray_plane(rays_out, cap, 2) + o, out.c); st_stream(ray_plane(rays_out, cap, 3) + o, out.d);
  }
}

// Scene-specialised instantiations (chosen at launch from DScene::spec_bits, device_scene.h SPEC_*): code a scene
// never runs still costs it registers and instruction-cache misses in this 70 KB kernel (c4: +3.7 %).
// Occupancy: the instantiations without the generic medium probes fit 64 registers with ~60 B of spills and
// run 8 blocks per SM (c4: 16.5 -> 16.2 ms per row); the others stay at 7 blocks / 72 registers.
// DEFER: items of the rare, heavy, textured classes (image / Perlin Lambertians: 0.8 % of c4's items, but 8 % of this
// kernel's warp instructions at 5 of 32 lanes, and 20 KB of its code) are not shaded here: their queue positions
// go to a list that k_wf_shade_rare works off densely, and this instantiation carries no texture code at all.
template <bool STATS, int SPEC, bool DEFER = false>
__global__ void __launch_bounds__(WF_SHADE_BLOCK, (SPEC & SPEC_GENERIC_MEDIA) ? WF_SHADE_MIN_BLOCKS : WF_SHADE_MIN_BLOCKS + 1) k_wf_shade(const __grid_constant__ DScene S, WFQueues Q,
                                                            const RayRec* __restrict__ rays_in,
                                                            RayRec* __restrict__ rays_out, float4* __restrict__ accum,
                                                            DStats* __restrict__ stats) {
  // (Sorting INDICES and letting the shading lane re-read its item from the queue -- 3 KB of shared memory per
  //  block instead of 11 KB, more L1 for the gathers -- was measured: 19.0 -> 20.5 ms per c4 row.)
  __shared__ ShadeItem items[WF_SHADE_BLOCK];
  __shared__ int class_count[NUM_CLASSES];
  const int tid = threadIdx.x, lane = tid & 31;
  const int i = blockIdx.x * WF_SHADE_BLOCK + tid;
  const int n = Q.c->n_in;
  if (blockIdx.x * WF_SHADE_BLOCK >= n) return;  // whole block idle (uniform)
  const Tables T = scene_tables(S);
  if (tid < NUM_CLASSES) class_count[tid] = 0;
  // (Asking the L2 for the tile a block ~1000 positions further on will stream -- cp.async.bulk.prefetch.L2 of
  //  its five planes by one thread -- was measured: 19.0 -> 19.3 ms per c4 row.  The DRAM latency of these
  //  loads is not what this kernel waits for.)
  // ---- 1. load + medium events + class -----------------------------------------------------------
  ShadeItem it;
  int cls = -1;
  const int cap = Q.capacity;
  if (i < n) {
    it.a = ld_stream(ray_plane(rays_in, cap, 0) + i); it.b = ld_stream(ray_plane(rays_in, cap, 1) + i);
    it.c = ld_stream(ray_plane(rays_in, cap, 2) + i); it.d = ld_stream(ray_plane(rays_in, cap, 3) + i);
    const uint4 h = ld_stream(reinterpret_cast<const uint4*>(Q.hits + i));
    it.t = __hiloint2double((int)h.y, (int)h.x); it.id = (int)h.z; it.info_x = (int)h.w;
  }
  __syncthreads();  // counters zeroed
  if (i < n) cls = wf_resolve<SPEC>(S, T, it.a, it.b, it.c, it.d, it.t, it.id, it.info_x);
  if (DEFER) {
    // a deferred item is a surface hit no medium event came before, so its hit record already says everything
    const bool defer = (cls == CLS_LAMBERT_TEX || cls == CLS_NOISE) && it.id >= 0;
    const unsigned dm = __ballot_sync(0xFFFFFFFFu, defer);
    if (dm) {
      const int leader = __ffs(dm) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&Q.c->n_deferred, __popc(dm));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (defer) {
        Q.deferred[base + __popc(dm & ((1u << lane) - 1u))] = i;
        cls = -1;
      }
    }
  }
  // ---- 2. block-local counting sort by class ---------------------------------------------------------
  int dst = wf_class_slot(cls, lane, class_count);
  __syncthreads();
  if (cls >= 0) items[dst + wf_class_prefix(cls, class_count)] = it;
  __syncthreads();
  const int n_block = min(WF_SHADE_BLOCK, n - blockIdx.x * WF_SHADE_BLOCK);
  int n_sorted = n_block;
  if (DEFER) {
    n_sorted = 0;
#pragma unroll
    for (int k = 0; k < NUM_CLASSES; k++) n_sorted += class_count[k];
  }
  // ---- 3. shade the item at sorted position `tid` -----------------------------------------------------
  bool alive = false;
  RayRec out;
  DStats st = {0, 0, 0, 0, 0, 0};
  if (tid < n_sorted) {
    const ShadeItem me = items[tid];
    alive = wf_shade_item<STATS, SPEC>(S, T, me.a, me.b, me.c, me.d, me.t, me.id, me.info_x, accum, out, st);
  }
  wf_append(Q, rays_out, alive, out, lane);
  if (STATS) {
    if (st.nonfinite) atomicAdd(&stats->nonfinite, st.nonfinite);
    if (tid == 0 && S.n_media > 0) atomicAdd(&stats->medium_probes, (unsigned long long)S.n_media * (unsigned long long)n_block);
  }
}

// the deferred items of k_wf_shade<.., DEFER = true>: same shading (every feature compiled in), dense lanes
template <bool STATS>
__global__ void __launch_bounds__(WF_SHADE_BLOCK) k_wf_shade_rare(const __grid_constant__ DScene S, WFQueues Q,
                                                                   const RayRec* __restrict__ rays_in, RayRec* __restrict__ rays_out,
                                                                   float4* __restrict__ accum, DStats* __restrict__ stats) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int nd = Q.c->n_deferred, cap = Q.capacity;
  const Tables T = scene_tables(S);
  DStats st = {0, 0, 0, 0, 0, 0};
  for (int base = blockIdx.x * WF_SHADE_BLOCK; base < nd; base += gridDim.x * WF_SHADE_BLOCK) {  // uniform per block
    bool alive = false;
    RayRec out;
    if (base + tid < nd) {
      const int j = Q.deferred[base + tid];
      const uint4 a = __ldg(ray_plane(rays_in, cap, 0) + j), b = __ldg(ray_plane(rays_in, cap, 1) + j);
      const uint4 c = __ldg(ray_plane(rays_in, cap, 2) + j), d = __ldg(ray_plane(rays_in, cap, 3) + j);
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(Q.hits + j));
      alive = wf_shade_item<STATS, SPEC_ALL>(S, T, a, b, c, d, __hiloint2double((int)h.y, (int)h.x), (int)h.z, (int)h.w, accum, out, st);
    }
    wf_append(Q, rays_out, alive, out, lane);
  }
  if (STATS && st.nonfinite) atomicAdd(&stats->nonfinite, st.nonfinite);
}

// ------------------------------------------------------------------------------------------------
// shade, persistent + TMA-staged variant.  Same three steps, but a block walks the queue in tiles of
// WF_SHADE_BLOCK slots and the NEXT tile's five 2 KB planes (ray words a-d + hit records) are already in
// flight -- cp.async.bulk into a second shared-memory stage, completion on an mbarrier -- while the
// current tile is resolved, sorted and shaded, so the DRAM latency of the queue records is off the critical
// path.  The sort permutes INDICES (order[]) and the shading lane reads its item straight from the stage,
// so the 80-byte items are no longer copied through shared memory either.
// MEASURED on c4 (60 GPU tests green with it): 23.6 ms per row against 19.0 ms for k_wf_shade at 7 blocks/SM.
// The two 10 KB stages per block take 165 KB of the SM's 228 KB away from the L1, which this kernel's
// primitive / material / texture gathers live on, and the latency it hides was not the one that
// mattered (see the L2-prefetch note in k_wf_shade).  Opt-in: RTB_WF_SHADE_TMA=1.
// ------------------------------------------------------------------------------------------------
struct alignas(128) ShadeStage {
  uint4 a[WF_SHADE_BLOCK], b[WF_SHADE_BLOCK], c[WF_SHADE_BLOCK], d[WF_SHADE_BLOCK], h[WF_SHADE_BLOCK];
};
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#ifndef WF_SHADE_TMA_MIN_BLOCKS
#define WF_SHADE_TMA_MIN_BLOCKS 6
#endif
template <bool STATS>
__global__ void __launch_bounds__(WF_SHADE_BLOCK, WF_SHADE_TMA_MIN_BLOCKS) k_wf_shade_tma(const __grid_constant__ DScene S, WFQueues Q,
                                                                const RayRec* __restrict__ rays_in,
                                                                RayRec* __restrict__ rays_out, float4* __restrict__ accum,
                                                                DStats* __restrict__ stats) {
  __shared__ ShadeStage stage[2];
  __shared__ alignas(8) unsigned long long full[2];
  __shared__ double res_t[WF_SHADE_BLOCK];
  __shared__ int res_id[WF_SHADE_BLOCK];
  __shared__ int order[WF_SHADE_BLOCK];
  __shared__ int class_count[NUM_CLASSES];
  const int tid = threadIdx.x, lane = tid & 31;
  const int n = Q.c->n_in;
  const int n_tiles = (n + WF_SHADE_BLOCK - 1) / WF_SHADE_BLOCK;
  if ((int)blockIdx.x >= n_tiles) return;
  const Tables T = scene_tables(S);
  const int cap = Q.capacity;
  // one thread arms the stage's mbarrier with the byte count and issues the five bulk copies of a tile
  auto issue = [&](int tile, int s) {
    const int base = tile * WF_SHADE_BLOCK;
    const unsigned bytes = (unsigned)min(WF_SHADE_BLOCK, n - base) * 16u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this stage are done (barrier)
    mbar_expect_tx(&full[s], 5u * bytes);
    bulk_g2s(stage[s].a, ray_plane(rays_in, cap, 0) + base, bytes, &full[s]);
    bulk_g2s(stage[s].b, ray_plane(rays_in, cap, 1) + base, bytes, &full[s]);
    bulk_g2s(stage[s].c, ray_plane(rays_in, cap, 2) + base, bytes, &full[s]);
    bulk_g2s(stage[s].d, ray_plane(rays_in, cap, 3) + base, bytes, &full[s]);
    bulk_g2s(stage[s].h, Q.hits + base, bytes, &full[s]);
  };
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < NUM_CLASSES) class_count[tid] = 0;
  __syncthreads();
  if (tid == 0) issue((int)blockIdx.x, 0);
  DStats st = {0, 0, 0, 0, 0, 0};
  unsigned long long probes = 0;
  for (int tile = (int)blockIdx.x, k = 0; tile < n_tiles; tile += (int)gridDim.x, k++) {
    const int s = k & 1;
    const int next = tile + (int)gridDim.x;
    if (tid == 0 && next < n_tiles) issue(next, s ^ 1);  // the other stage was released by the barrier that ended the last tile
    mbar_wait(&full[s], (unsigned)(k >> 1) & 1u);
    const int n_block = min(WF_SHADE_BLOCK, n - tile * WF_SHADE_BLOCK);
    // ---- 1. medium events + class --------------------------------------------------------------------
    int cls = -1;
    if (tid < n_block) {
      const uint4 h = stage[s].h[tid];
      double t = __hiloint2double((int)h.y, (int)h.x);
      int id = (int)h.z;
      cls = wf_resolve<SPEC_ALL>(S, T, stage[s].a[tid], stage[s].b[tid], stage[s].c[tid], stage[s].d[tid], t, id, (int)h.w);
      res_t[tid] = t;
      res_id[tid] = id;
    }
    // ---- 2. counting sort of the tile's indices by class ------------------------------------------------
    const int dst = wf_class_slot(cls, lane, class_count);
    __syncthreads();
    if (cls >= 0) order[dst + wf_class_prefix(cls, class_count)] = tid;
    __syncthreads();
    if (tid < NUM_CLASSES) class_count[tid] = 0;  // dead until the next tile's sort; ordered by the barrier below
    // ---- 3. shade the item at sorted position `tid`, read straight from the stage ---------------------------
    bool alive = false;
    RayRec out;
    if (tid < n_block) {
      const int src = order[tid];
      const uint4 h = stage[s].h[src];
      alive = wf_shade_item<STATS, SPEC_ALL>(S, T, stage[s].a[src], stage[s].b[src], stage[s].c[src], stage[s].d[src], res_t[src], res_id[src],
                                   (int)h.w, accum, out, st);
    }
    wf_append(Q, rays_out, alive, out, lane);
    if (STATS && tid == 0) probes += (unsigned long long)S.n_media * (unsigned long long)n_block;
    __syncthreads();  // stage s, order[], res_*[] may be overwritten from here on
  }
  if (STATS) {
    if (st.nonfinite) atomicAdd(&stats->nonfinite, st.nonfinite);
    if (tid == 0 && probes) atomicAdd(&stats->medium_probes, probes);
  }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
template <int SPEC>
static void launch_shade_spec(const DScene& S, const WFQueues& Q, const RayRec* in, RayRec* out, float4* d_accum, DStats* d_stats,
                              unsigned blocks, cudaStream_t st) {
  k_wf_shade<false, SPEC><<<blocks, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
}
template <int SPEC>
static void launch_shade_defer(const DScene& S, const WFQueues& Q, const RayRec* in, RayRec* out, float4* d_accum, DStats* d_stats,
                               unsigned blocks, int sms, cudaStream_t st) {
  k_wf_shade<false, SPEC, true><<<blocks, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
  const unsigned rare = std::max(1u, std::min(blocks / 16u, (unsigned)(sms * 4)));
  k_wf_shade_rare<false><<<rare, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
}
// returns the number of kernels launched
static int launch_shade(const DScene& S, const WFQueues& Q, const RayRec* in, RayRec* out, float4* d_accum, DStats* d_stats,
                        bool collect_stats, unsigned blocks, bool defer_rare, int sms, cudaStream_t st) {
  constexpr int kTex = SPEC_TEXTURES | SPEC_SPHERE_UV | SPEC_QUAD_UV;
  if (!collect_stats && defer_rare && S.defer_ok && (S.spec_bits & SPEC_TEXTURES)) {
    const int need = S.spec_bits & SPEC_ALL & ~kTex;
    if (need == SPEC_MEDIA) { launch_shade_defer<SPEC_MEDIA>(S, Q, in, out, d_accum, d_stats, blocks, sms, st); return 2; }
    if (need == (SPEC_MEDIA | SPEC_LIGHTS)) { launch_shade_defer<SPEC_MEDIA | SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, sms, st); return 2; }
    if (need == 0) { launch_shade_defer<0>(S, Q, in, out, d_accum, d_stats, blocks, sms, st); return 2; }
  }
  if (collect_stats) {  // the counted passes are not timed: one generic instantiation
    k_wf_shade<true, SPEC_ALL><<<blocks, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
    return 1;
  }
  // the smallest instantiation whose features cover the scene's (a superset is always correct)
  constexpr int kC4 = SPEC_MEDIA | SPEC_SPHERE_UV | SPEC_TEXTURES, kC3 = SPEC_MEDIA | SPEC_BOXSCAN | SPEC_GENERIC_MEDIA;
  const int need = S.spec_bits & SPEC_ALL;
  auto covers = [need](int spec) { return (spec & need) == need; };
  if (covers(0)) launch_shade_spec<0>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(SPEC_LIGHTS)) launch_shade_spec<SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC4)) launch_shade_spec<kC4>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC3)) launch_shade_spec<kC3>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC4 | SPEC_LIGHTS)) launch_shade_spec<kC4 | SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC3 | SPEC_LIGHTS)) launch_shade_spec<kC3 | SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else launch_shade_spec<SPEC_ALL>(S, Q, in, out, d_accum, d_stats, blocks, st);
  return 1;
}

// accum.w += number of strata rendered, for every pixel (what one atomicAdd(+1) per finished path would sum to)
__global__ void k_wf_add_count(float4* __restrict__ accum, int n_pixels, float n_strata) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pixels) accum[i].w += n_strata;
}

__global__ void k_wf_sum_segments(WFQueues Q0, WFQueues Q1, WFQueues Q2, WFQueues Q3, int n, DStats* stats) {
  const WFQueues* q[4] = {&Q0, &Q1, &Q2, &Q3};
  unsigned long long s = 0;
  for (int k = 0; k < n; k++) s += q[k]->c->segments;
  stats->segments = s;
}

static int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}

cudaError_t wavefront_context_create(WavefrontContext* ctx) {
  *ctx = WavefrontContext{};
  cudaError_t e = cudaMallocHost(&ctx->host_counters, WF_MAX_SUB * sizeof(WFCounters));
  if (e != cudaSuccess) return e;
  int dev = 0;
  cudaGetDevice(&dev);
  ctx->sms = 148;
  cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, dev);
  ctx->extend_blocks_per_sm[0] = ctx->extend_blocks_per_sm[1] = 4;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->extend_blocks_per_sm[0], k_wf_extend<false, NODES_BVH2>, WF_EXTEND_BLOCK, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->extend_blocks_per_sm[1], k_wf_extend<true, NODES_BVH2>, WF_EXTEND_BLOCK, 0);
  {  // the other node-format instantiations may differ by a few registers: take the smallest residency
    int r[4] = {0, 0, 0, 0};
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r[0], k_wf_extend<false, NODES_Q>, WF_EXTEND_BLOCK, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r[1], k_wf_extend<true, NODES_Q>, WF_EXTEND_BLOCK, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r[2], k_wf_extend<false, NODES_BVH4>, WF_EXTEND_BLOCK, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r[3], k_wf_extend<true, NODES_BVH4>, WF_EXTEND_BLOCK, 0);
    int m[2] = {0, 0};
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m[0], k_wf_extend<false, NODES_BVH2_MULTI>, WF_EXTEND_BLOCK, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m[1], k_wf_extend<true, NODES_BVH2_MULTI>, WF_EXTEND_BLOCK, 0);
    for (int k = 0; k < 2; k++)
      if (m[k] > 0 && m[k] < ctx->extend_blocks_per_sm[k]) ctx->extend_blocks_per_sm[k] = m[k];
    for (int k = 0; k < 4; k++)
      if (r[k] > 0 && r[k] < ctx->extend_blocks_per_sm[k & 1]) ctx->extend_blocks_per_sm[k & 1] = r[k];
  }
  ctx->pool_blocks_per_sm[0] = ctx->pool_blocks_per_sm[1] = 4;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->pool_blocks_per_sm[0], k_wf_extend_pool<false>, WF_POOL_BLOCK, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->pool_blocks_per_sm[1], k_wf_extend_pool<true>, WF_POOL_BLOCK, 0);
  ctx->defer_rare = env_int("RTB_WF_DEFER_RARE", 1, 0, 1);  // measured on c4: 16.0 -> 15.4 ms shade per row
  ctx->shade_tma = env_int("RTB_WF_SHADE_TMA", 0, 0, 1);
  ctx->shade_tma_blocks_per_sm = 4;
  {
    int a = 0, b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_wf_shade_tma<false>, WF_SHADE_BLOCK, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_wf_shade_tma<true>, WF_SHADE_BLOCK, 0);
    if (a > 0 && b > 0) ctx->shade_tma_blocks_per_sm = std::min(a, b);
    ctx->shade_tma_blocks_per_sm = env_int("RTB_WF_SHADE_TMA_BLOCKS", ctx->shade_tma_blocks_per_sm, 1, 32);
  }
  ctx->extend_kind = env_int("RTB_WF_EXTEND_POOL", 0, 0, 1);  // measured slower on c4: profiles/r01_pool_extend.txt
  // Sub-pipelines: independent slices of the stratum range on their own streams (RTB_WF_STREAMS).
  // Measured on c4 (profiles/r01_streams_sweep.txt): 2-4 streams are 1-9 % SLOWER than one -- the
  // persistent extend grid owns the register file, so the other stream's kernels queue behind it.
  // Default 1; the path stays for scenes whose stages might complement each other better.
  ctx->n_sub = env_int("RTB_WF_STREAMS", 1, 1, WF_MAX_SUB);
  for (int k = 0; k < ctx->n_sub; k++) {
    if ((e = cudaStreamCreateWithFlags(&ctx->streams[k], cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  if ((e = cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming)) != cudaSuccess) return e;
  return cudaGetLastError();
}

void wavefront_context_destroy(WavefrontContext* ctx) {
  if (ctx->host_counters) cudaFreeHost(ctx->host_counters);
  for (int k = 0; k < WF_MAX_SUB; k++) {
    if (ctx->streams[k]) cudaStreamDestroy(ctx->streams[k]);
    if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
  }
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  *ctx = WavefrontContext{};
}

cudaError_t launch_render_wavefront(const DScene& S, const WavefrontContext& ctx, int64_t s_begin, int64_t s_end,
                                    float4* d_accum, DStats* d_stats, bool collect_stats, void* d_workspace,
                                    size_t workspace_bytes, int64_t capacity, cudaStream_t stream, int* launches) {
  if (workspace_bytes < wavefront_workspace_bytes(S, capacity)) return cudaErrorInvalidValue;
  // RTB_WF_PROFILE=1: one pipeline, per-stage CUDA-event totals on stderr (analysis runs only)
  static const bool profile = getenv("RTB_WF_PROFILE") != nullptr;
  const long long n_strata = (long long)(s_end - s_begin);
  int K = profile ? 1 : ctx.n_sub;
  if (n_strata < K) K = (int)n_strata;
  if (K < 1) return cudaSuccess;
  const int64_t cap = (capacity / K) & ~(int64_t)255;
  const unsigned long long tiles = (unsigned long long)((S.cam.width + 7) / 8) * ((S.cam.height + 3) / 4);
  const int occ = ctx.extend_blocks_per_sm[collect_stats ? 1 : 0] < 1 ? 1 : ctx.extend_blocks_per_sm[collect_stats ? 1 : 0];
  const int per_sm = env_int("RTB_WF_EXTEND_BLOCKS", (occ + K - 1) / K, 1, 32);
  const long long extend_grid_full = (long long)ctx.sms * per_sm;
  // pool extend: needs the stacks to fit WF_POOL_STACK; one sub-pipeline only (it owns the scratch)
  const bool use_pool = ctx.extend_kind == 1 && K == 1 && S.bvh_depth + 2 <= WF_POOL_STACK && ctx.sms <= 160;
  int pool_per_sm = ctx.pool_blocks_per_sm[collect_stats ? 1 : 0];
  pool_per_sm = env_int("RTB_WF_POOL_BLOCKS", pool_per_sm < 1 ? 1 : pool_per_sm, 1, WF_POOL_MAX_BLOCKS_PER_SM);
  if (pool_per_sm > WF_POOL_MAX_BLOCKS_PER_SM) pool_per_sm = WF_POOL_MAX_BLOCKS_PER_SM;
  const long long pool_grid_full = (long long)ctx.sms * pool_per_sm;
  int* const pool_scratch = reinterpret_cast<int*>(static_cast<char*>(d_workspace) + workspace_bytes - pool_scratch_bytes(160));

  struct Sub { WFQueues Q; RayRec* in; RayRec* out; cudaStream_t st; long long s0; bool active; };
  Sub sub[WF_MAX_SUB];
  const size_t sub_bytes = queue_bytes(cap);
  if ((size_t)K * sub_bytes + pool_scratch_bytes(160) > workspace_bytes) return cudaErrorInvalidValue;
  cudaError_t e = cudaEventRecord(ctx.ev_start, stream);
  if (e != cudaSuccess) return e;
  int n_launch = 0;
  for (int k = 0; k < K; k++) {
    Sub& u = sub[k];
    u.Q = carve(static_cast<char*>(d_workspace) + (size_t)k * sub_bytes, cap);
    u.in = u.Q.rays_a;
    u.out = u.Q.rays_b;
    u.st = profile ? stream : ctx.streams[k];
    u.active = true;
    const long long lo = n_strata * k / K, hi = n_strata * (k + 1) / K;
    u.s0 = (long long)s_begin + lo;
    if (!profile && (e = cudaStreamWaitEvent(u.st, ctx.ev_start, 0)) != cudaSuccess) return e;
    k_wf_init<<<1, 1, 0, u.st>>>(u.Q, tiles * 32ull * (unsigned long long)(hi - lo));
    n_launch++;
  }
  WFCounters* h_c = static_cast<WFCounters*>(ctx.host_counters);  // pinned mirrors for the (sparse) host polls
  cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
  double stage_ms[3] = {0., 0., 0.};
  if (profile)
    for (auto& ev : pe) cudaEventCreate(&ev);
  int poll_every = profile ? 1 : 8;
  int n_active = K;
  long long iters = 0;
  // Upper bound of the rays of the coming iterations, known to the host only at polls: while paths are
  // still being started the queue is full; once every path has started it can only shrink, so the last
  // polled count bounds all later iterations and the launches of the decaying tail are sized to it.
  long long bound[WF_MAX_SUB];
  for (int k = 0; k < K; k++) bound[k] = cap;
  for (long long iter = 0; n_active > 0; iter++) {
    for (int k = 0; k < K; k++) {
      Sub& u = sub[k];
      if (!u.active) continue;
      const unsigned gen_blocks = (unsigned)((bound[k] + 255) / 256);
      const unsigned shade_blocks = (unsigned)((bound[k] + WF_SHADE_BLOCK - 1) / WF_SHADE_BLOCK);
      const unsigned extend_grid = (unsigned)std::min<long long>(extend_grid_full, (bound[k] + WF_EXTEND_BLOCK - 1) / WF_EXTEND_BLOCK);
      const unsigned pool_grid = (unsigned)std::min<long long>(pool_grid_full, (bound[k] + 2 * WF_POOL_BLOCK - 1) / (2 * WF_POOL_BLOCK));
      // top the out queue up (first iteration: fill it), then it becomes this iteration's in queue
      if (profile) cudaEventRecord(pe[0], u.st);
      k_wf_generate<<<gen_blocks, 256, 0, u.st>>>(S, u.Q, u.s0, u.out);
      k_wf_advance<<<1, 1, 0, u.st>>>(u.Q);
      if (profile) cudaEventRecord(pe[1], u.st);
      { RayRec* t = u.in; u.in = u.out; u.out = t; }
      if (use_pool) {
        if (collect_stats) k_wf_extend_pool<true><<<pool_grid, WF_POOL_BLOCK, 0, u.st>>>(S, u.Q, u.in, pool_scratch, d_stats);
        else k_wf_extend_pool<false><<<pool_grid, WF_POOL_BLOCK, 0, u.st>>>(S, u.Q, u.in, pool_scratch, d_stats);
      } else {
        if (S.multi_leaf) {  // leaves of several primitives: the generic leaf loop, fp32 BVH2 nodes
          if (collect_stats) k_wf_extend<true, NODES_BVH2_MULTI><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
          else k_wf_extend<false, NODES_BVH2_MULTI><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
        } else if (S.use_qnodes) {
          if (collect_stats) k_wf_extend<true, NODES_Q><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
          else k_wf_extend<false, NODES_Q><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
        } else if (S.use_bvh4) {
          if (collect_stats) k_wf_extend<true, NODES_BVH4><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
          else k_wf_extend<false, NODES_BVH4><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
        } else {
          if (collect_stats) k_wf_extend<true, NODES_BVH2><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
          else k_wf_extend<false, NODES_BVH2><<<extend_grid, WF_EXTEND_BLOCK, 0, u.st>>>(S, u.Q, u.in, d_stats);
        }
      }
      if (profile) cudaEventRecord(pe[2], u.st);
      if (ctx.shade_tma) {  // persistent blocks, tiles strided over the grid
        const unsigned grid = (unsigned)std::min<long long>(shade_blocks, (long long)ctx.sms * ctx.shade_tma_blocks_per_sm);
        if (collect_stats) k_wf_shade_tma<true><<<grid, WF_SHADE_BLOCK, 0, u.st>>>(S, u.Q, u.in, u.out, d_accum, d_stats);
        else k_wf_shade_tma<false><<<grid, WF_SHADE_BLOCK, 0, u.st>>>(S, u.Q, u.in, u.out, d_accum, d_stats);
      } else {
        n_launch += launch_shade(S, u.Q, u.in, u.out, d_accum, d_stats, collect_stats, shade_blocks, ctx.defer_rare != 0, ctx.sms, u.st) - 1;
      }
      if (profile) {
        cudaEventRecord(pe[3], u.st);
        cudaEventSynchronize(pe[3]);
        float it_ms[3];
        for (int j = 0; j < 3; j++) { it_ms[j] = 0.f; cudaEventElapsedTime(&it_ms[j], pe[j], pe[j + 1]); stage_ms[j] += it_ms[j]; }
        static const bool per_iter = getenv("RTB_WF_PROFILE") && atoi(getenv("RTB_WF_PROFILE")) >= 2;
        if (per_iter) fprintf(stderr, "[rtb iter] %lld extend %.3f shade %.3f\n", iter, it_ms[1], it_ms[2]);
      }
      n_launch += 4;
    }
    iters++;
    if ((iter % poll_every) == poll_every - 1) {
      for (int k = 0; k < K; k++)
        if (sub[k].active && (e = cudaMemcpyAsync(h_c + k, sub[k].Q.c, sizeof(WFCounters), cudaMemcpyDeviceToHost, sub[k].st)) != cudaSuccess) return e;
      for (int k = 0; k < K; k++) {
        if (!sub[k].active) continue;
        if ((e = cudaStreamSynchronize(sub[k].st)) != cudaSuccess) return e;
        if (h_c[k].next_path >= h_c[k].total_paths) {
          if (h_c[k].n_out == 0) { sub[k].active = false; n_active--; }
          bound[k] = std::max<long long>(h_c[k].n_out, 1);  // every path has started: the queue only shrinks
          if (!profile) poll_every = bound[k] < cap / 4 ? 2 : 4;
        }
      }
    }
  }
  if (profile) {
    fprintf(stderr, "[rtb wavefront] iterations %lld  segments %llu  generate %.2f ms  extend %.2f ms  shade %.2f ms\n", iters,
            (unsigned long long)h_c[0].segments, stage_ms[0], stage_ms[1], stage_ms[2]);
    for (auto& ev : pe) cudaEventDestroy(ev);
  } else {
    for (int k = 0; k < K; k++) {
      if ((e = cudaEventRecord(ctx.ev_done[k], sub[k].st)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(stream, ctx.ev_done[k], 0)) != cudaSuccess) return e;
    }
  }
  {
    const int n_pixels = S.cam.width * S.cam.height;
    k_wf_add_count<<<(n_pixels + 255) / 256, 256, 0, stream>>>(d_accum, n_pixels, (float)n_strata);
    n_launch++;
  }
  if (collect_stats) {
    k_wf_sum_segments<<<1, 1, 0, stream>>>(sub[0].Q, sub[K > 1 ? 1 : 0].Q, sub[K > 2 ? 2 : 0].Q, sub[K > 3 ? 3 : 0].Q, K, d_stats);
    n_launch++;
  }
  if (launches) *launches += n_launch;
  return cudaGetLastError();
}

}  // namespace rtb
