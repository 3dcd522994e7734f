// wavefront.cu -- the wavefront pipeline of the hot path (the product pipeline; the megakernel in
// kernels.cu is the arm for small jobs and the A/B reference).
//
// One iteration advances every in-flight path by one segment:
//
//   k_wf_generate      tops the ray queue up with new (pixel, stratum) rays and   get_ray  render.rs:218-249
//                      advances the queue counters (last block done)
//   k_wf_extend        BVH traversal: persistent warps, dynamic ray fetch,         HittableList::hit / BvhNode::hit
//                      speculative while-while descent.  Leaf primitives are only    hittable.rs:88-109, 216-236
//                      CLASSIFIED (conservative fp32, rtb_device.cuh prefilter_*);
//                      the <= 2 that can still be the closest hit go to the hit record
//   k_wf_extend_exact  the rare rays with more live candidates than slots, re-traced with the exact tests
//   k_wf_shade         exact f64 reference-order tests of the candidates (dense     Sphere::hit / Quad::hit  object.rs:145-184, 453-490
//                      lanes), constant-medium events, block-local sort by shading   ConstantMedium::hit  constant_medium.rs:41-95
//                      class, emit / scatter / mixture-pdf sample; survivors are     ray_color  render.rs:271-297
//                      appended densely to the next queue, finished paths accumulate into the image
//   k_wf_finish        once few paths are left, one kernel runs them to their end (no more ~50 us iterations)
//
// Data layout in HBM (DESIGN.md "Queues"): two dense ray queues, each FOUR uint4 PLANES of `capacity`
// entries (SoA of a 64-byte record) in QUEUE ORDER -- no slot indirection, every stage streams them and a
// warp's 32 consecutive slots are 512 contiguous bytes per plane.  Two encodings share the record:
//   secondary (bounce >= 1): origin f64 x3 | direction f32 x3, time f32 | throughput f32 x3 | pixel, stratum, bounce
//   primary   (bounce == 0): origin f64 x3 | direction f64 x3 (throughput is 1)  | time f32 | pixel, stratum, 0
// so that camera rays reach the exact tests with the reference's f64 directions (get_ray, render.rs:221-232).
// Survivors come first in a queue, fresh primaries after them: position >= n_surv <=> primary.
// One 8-byte candidate record {leaf ref, leaf ref} per queue position; all counters live on the device, the
// host only polls "paths left" every few iterations.
//
// Accumulation: 64-bit fixed-point sums (2^-32 units) per channel, integer atomics -- order-independent, so a
// render is bit-reproducible run to run and across any split of the stratum range over calls or GPUs.
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
  int n_surv;         // queue positions below this hold survivors (secondary encoding), the rest fresh primaries
  int n_overflow;     // rays of this iteration whose candidates overflowed (re-traced by k_wf_extend_exact)
  int gen_done;       // blocks of k_wf_generate that have finished (the last one advances the counters)
  int pad0;
  unsigned long long next_path;    // camera paths started so far
  unsigned long long total_paths;  // to start in this render call (padded tiles included)
  unsigned long long segments;     // sum of n_in over iterations
  unsigned long long overflows;    // sum of n_overflow over iterations
};

struct RayRec { uint4 a, b, c, d; };  // the four words of one ray (see pack/unpack); stored as planes, see ray_plane
struct alignas(16) HitRec { double t; int id; int info_x; };  // resolved hit of a DEFERRED item (k_wf_shade_rare)

struct WFQueues {
  RayRec* rays_a;
  RayRec* rays_b;
  int2* cands;        // per queue position: the candidate leaf references k_wf_extend leaves for the shade stage
  HitRec* hits;       // per queue position, written only for deferred items
  int* deferred;      // queue positions of deferred items (see k_wf_shade_rare)
  int* overflow;      // queue positions of rays to re-trace exactly
  WFCounters* c;
  int capacity;
};

struct PathRec {  // unpacked RayRec
  double ox, oy, oz;
  double dx, dy, dz;  // f64: exact for primaries, the widened fp32 values for secondaries
  float time;
  float bx, by, bz;
  uint32_t pixel, sample, bounce;
};

__device__ __forceinline__ double bits_to_double(unsigned lo, unsigned hi) { return __hiloint2double((int)hi, (int)lo); }

__device__ __forceinline__ RayRec pack_secondary(const PathRec& p) {
  RayRec r;
  r.a.x = (unsigned)__double2loint(p.ox); r.a.y = (unsigned)__double2hiint(p.ox);
  r.a.z = (unsigned)__double2loint(p.oy); r.a.w = (unsigned)__double2hiint(p.oy);
  r.b.x = (unsigned)__double2loint(p.oz); r.b.y = (unsigned)__double2hiint(p.oz);
  r.b.z = __float_as_uint((float)p.dx); r.b.w = __float_as_uint((float)p.dy);
  r.c.x = __float_as_uint((float)p.dz); r.c.y = __float_as_uint(p.time);
  r.c.z = __float_as_uint(p.bx); r.c.w = __float_as_uint(p.by);
  r.d.x = __float_as_uint(p.bz); r.d.y = p.pixel; r.d.z = p.sample; r.d.w = p.bounce;
  return r;
}
__device__ __forceinline__ RayRec pack_primary(const PathRec& p) {
  RayRec r;
  r.a.x = (unsigned)__double2loint(p.ox); r.a.y = (unsigned)__double2hiint(p.ox);
  r.a.z = (unsigned)__double2loint(p.oy); r.a.w = (unsigned)__double2hiint(p.oy);
  r.b.x = (unsigned)__double2loint(p.oz); r.b.y = (unsigned)__double2hiint(p.oz);
  r.b.z = (unsigned)__double2loint(p.dx); r.b.w = (unsigned)__double2hiint(p.dx);
  r.c.x = (unsigned)__double2loint(p.dy); r.c.y = (unsigned)__double2hiint(p.dy);
  r.c.z = (unsigned)__double2loint(p.dz); r.c.w = (unsigned)__double2hiint(p.dz);
  r.d.x = __float_as_uint(p.time); r.d.y = p.pixel; r.d.z = p.sample; r.d.w = 0u;
  return r;
}
__device__ __forceinline__ void unpack(const uint4& a, const uint4& b, const uint4& c, const uint4& d, PathRec& p) {
  p.ox = bits_to_double(a.x, a.y); p.oy = bits_to_double(a.z, a.w); p.oz = bits_to_double(b.x, b.y);
  p.pixel = d.y; p.sample = d.z; p.bounce = d.w;
  if (d.w == 0u) {
    p.dx = bits_to_double(b.z, b.w); p.dy = bits_to_double(c.x, c.y); p.dz = bits_to_double(c.z, c.w);
    p.time = __uint_as_float(d.x);
    p.bx = p.by = p.bz = 1.f;
  } else {
    p.dx = (double)__uint_as_float(b.z); p.dy = (double)__uint_as_float(b.w); p.dz = (double)__uint_as_float(c.x);
    p.time = __uint_as_float(c.y);
    p.bx = __uint_as_float(c.z); p.by = __uint_as_float(c.w); p.bz = __uint_as_float(d.x);
  }
}
__device__ __forceinline__ Ray to_ray(const PathRec& p) {
  Ray r;
  r.ox = p.ox; r.oy = p.oy; r.oz = p.oz;
  r.dx = p.dx; r.dy = p.dy; r.dz = p.dz; r.time = (double)p.time;
  return r;
}

constexpr uint32_t PADDING_PIXEL = 0xFFFFFFFFu;  // inert lane of a border tile

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
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) { __stcs(p, v); }
__device__ __forceinline__ int2 ld_stream(const int2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(int2* p, const int2& v) { __stcs(p, v); }

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t wavefront_workspace_bytes(const DScene&, int64_t n) {
  return 2 * align_up(n * sizeof(RayRec)) + align_up(n * sizeof(int2)) + align_up(n * sizeof(HitRec)) + 2 * align_up(n * sizeof(int)) +
         align_up(sizeof(WFCounters)) + 4096;
}

size_t wavefront_counters_bytes() { return sizeof(WFCounters); }

static WFQueues carve(void* ws, int64_t n) {
  char* p = static_cast<char*>(ws);
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes); return r; };
  WFQueues q;
  q.rays_a = reinterpret_cast<RayRec*>(take(n * sizeof(RayRec)));
  q.rays_b = reinterpret_cast<RayRec*>(take(n * sizeof(RayRec)));
  q.cands = reinterpret_cast<int2*>(take(n * sizeof(int2)));
  q.hits = reinterpret_cast<HitRec*>(take(n * sizeof(HitRec)));
  q.deferred = reinterpret_cast<int*>(take(n * sizeof(int)));
  q.overflow = reinterpret_cast<int*>(take(n * sizeof(int)));
  q.c = reinterpret_cast<WFCounters*>(take(sizeof(WFCounters)));
  q.capacity = (int)n;
  return q;
}

// ------------------------------------------------------------------------------------------------
// bookkeeping
// ------------------------------------------------------------------------------------------------
__global__ void k_wf_init(WFQueues Q, unsigned long long total_paths) {
  WFCounters z = {};
  z.total_paths = total_paths;
  *Q.c = z;
}

// after shade + generate: the out queue becomes the in queue of the next iteration (one thread)
__device__ __forceinline__ void wf_advance(const WFQueues& Q) {
  WFCounters* c = Q.c;
  const unsigned long long left = c->total_paths - c->next_path;
  const unsigned long long room = (unsigned long long)(Q.capacity - c->n_out);
  const unsigned long long gen = left < room ? left : room;
  c->next_path += gen;
  c->n_surv = c->n_out;
  c->n_in = c->n_out + (int)gen;
  c->segments += (unsigned long long)c->n_in;
  c->overflows += (unsigned long long)c->n_overflow;
  c->n_out = 0;
  c->extend_cursor = 0;
  c->n_deferred = 0;
  c->n_overflow = 0;
  c->gen_done = 0;
}

// ------------------------------------------------------------------------------------------------
// generate: path id -> (stratum, pixel).  Ids enumerate 8x4 pixel tiles padded to 32 lanes, so the
// 32 consecutive rays of one warp-fetch in extend are one coherent tile.  The block that finishes last
// advances the queue counters (what used to be a one-thread kernel of its own: ~5 us per iteration).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_generate(const __grid_constant__ DScene S, WFQueues Q, long long s_begin,
                                                      RayRec* __restrict__ out) {
  WFCounters* c = Q.c;
  const unsigned long long next_path = c->next_path;
  const unsigned long long left = c->total_paths - next_path;
  const int n_out = c->n_out;
  // grid-stride over the free slots: the grid is capped (one arrival atomic per block below: 131 K blocks arriving
  // on one address cost 0.14 ms per iteration when the grid was sized to the queue)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Q.capacity - n_out && (unsigned long long)i < left;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long pid = next_path + (unsigned long long)i;
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
      p.dx = ps.ray.dx; p.dy = ps.ray.dy; p.dz = ps.ray.dz; p.time = (float)ps.ray.time;
      p.pixel = pixel;
    } else {  // padding lane: dies in its first shade without touching the image
      p.ox = p.oy = p.oz = 0.;
      p.dx = p.dy = 0.; p.dz = 1.; p.time = 0.f;
      p.pixel = PADDING_PIXEL;
    }
    p.sample = sample;
    const RayRec rec = pack_primary(p);
    const int o = n_out + (int)i, cap = Q.capacity;
    st_stream(ray_plane(out, cap, 0) + o, rec.a); st_stream(ray_plane(out, cap, 1) + o, rec.b);
    st_stream(ray_plane(out, cap, 2) + o, rec.c); st_stream(ray_plane(out, cap, 3) + o, rec.d);
  }
  // every block has read the counters before it arrives here, so the last arrival may rewrite them
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&c->gen_done, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) wf_advance(Q);
}

// ------------------------------------------------------------------------------------------------
// extend: persistent warps, dynamic fetch, speculative while-while traversal (Aila & Laine 2009):
// a lane that reaches a leaf postpones it and keeps descending until every lane of the warp holds a
// leaf or has run out of nodes; then the leaves are handled together.  fp32 conservative slabs in
// fused form (t = plane * inv_d - o * inv_d).
// CAND = true (default): a leaf primitive is classified by the conservative fp32 test and the survivors
//   travel to the shade stage as candidates; the cull bound is the smallest upper t of a certain hit.
// CAND = false (RTB_OPT_EXACT_LEAVES, the round-1 kernel, kept as the A/B arm of the same binary): the f64
//   reference-order tests run here, at 5-8 of 32 lanes; the winner is written as the only candidate.
// ------------------------------------------------------------------------------------------------
constexpr int WF_EXTEND_BLOCK = 128;
#ifndef WF_FETCH_THRESHOLD_N
#define WF_FETCH_THRESHOLD_N 20  // measured on c4: 32 -> 26.9, 28 -> 26.8, 24 -> 26.7, 20 -> 26.5 ms extend per row
#endif
constexpr int WF_FETCH_THRESHOLD = WF_FETCH_THRESHOLD_N;  // refill when fewer than this many lanes hold a ray
constexpr int TRAV_DONE = 0x7FFFFFFF;
#ifndef WF_DENSE_LEAVES
// 1: the leaf references of a whole warp are classified through a shared-memory list, lane j taking entry j whoever owns it.
// Measured on c4 (profiles/r02_leaf_phase_experiments.txt): extend 356 ms per step against 230 -- a round holds only ~15 leaf
// references per warp, so the leaf code still runs at 6-9 lanes, and the list, ballots and barriers add 25 % instructions.
// So did sorting the two slots of every lane into one pass per kind (box / sphere): 324 against 221.  Kept as the A/B arm.
#define WF_DENSE_LEAVES 0
#endif
#ifndef WF_BREAK_LEFT
#define WF_BREAK_LEFT 16  // measured on c4: 0 -> 35.1, 12 -> 33.9, 16 -> 33.4, 20 -> 33.5, 24 -> 33.8 ms extend per step
#endif
#ifndef WF_BREAK_RELATIVE
#define WF_BREAK_RELATIVE 0
#endif
#ifndef WF_SYNC_AFTER_LEAVES
#define WF_SYNC_AFTER_LEAVES 1
#endif
#ifndef WF_EXTEND_MIN_BLOCKS_CAND
#define WF_EXTEND_MIN_BLOCKS_CAND 9  // 36 warps per SM at 56 registers (16 B of spills); measured on c4: 8 -> 264.4, 9 -> 255.0, 10 -> 301 ms extend per step
#endif
#ifndef WF_PREFETCH_AHEAD
#define WF_PREFETCH_AHEAD 0  // queue positions ahead of the fetch cursor to pull into L2 (0 = off)
#endif

// NODES: which form of the tree is traversed, chosen per scene by the builder --
//   NODES_BVH2        the 64-byte fp32 nodes;
//   NODES_Q           the 32-byte quantised nodes (rtb_device.cuh, slab_box_q; RTB_FLAG_QNODES);
//   NODES_BVH4        every other level collapsed (RTB_FLAG_BVH4): half the dependent steps per ray;
//   NODES_BVH2_MULTI  the BVH2 with leaves of several primitives (RTB_FLAG_BVH_LEAF4): the generic leaf loop compiled in;
//   NODES_BVH2_SMEM   the BVH2 with its first WF_SMEM_NODES nodes staged in shared memory (RTB_OPT_SMEM_TOP, A/B arm).
enum : int { NODES_BVH2 = 0, NODES_Q = 1, NODES_BVH4 = 2, NODES_BVH2_MULTI = 3, NODES_BVH2_SMEM = 4 };
constexpr int WF_SMEM_NODES = 127;  // 7 complete levels of a balanced tree, 8 KB per block

struct alignas(16) LeafRay { double ox, oy, oz; float dx, dy, dz, time; float pad, pad2; };  // 48 B: a lane's ray as the dense leaf phase reads it

template <bool STATS, int NODES, bool CAND>
__global__ void __launch_bounds__(WF_EXTEND_BLOCK, CAND ? WF_EXTEND_MIN_BLOCKS_CAND : 8) k_wf_extend(const __grid_constant__ DScene S, WFQueues Q,
                                                               const RayRec* __restrict__ rays_in,
                                                               DStats* __restrict__ stats) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  __shared__ float4 top_nodes[NODES == NODES_BVH2_SMEM ? 4 * WF_SMEM_NODES : 1];
  const int n_top = NODES == NODES_BVH2_SMEM ? min(S.n_nodes, WF_SMEM_NODES) : 0;
  if (NODES == NODES_BVH2_SMEM) {
    for (int k = threadIdx.x; k < 4 * n_top; k += blockDim.x) top_nodes[k] = __ldg(S.nodes + k);
    __syncthreads();
  }
#if WF_DENSE_LEAVES
  __shared__ LeafRay leaf_rays[CAND ? WF_EXTEND_BLOCK : 1];
  __shared__ int4 leaf_list[CAND ? WF_EXTEND_BLOCK / 32 : 1][64];
#endif
  const int n = Q.c->n_in, n_surv = Q.c->n_surv;
  unsigned long long st_nodes = 0, st_prims = 0;
  bool have = false, exhausted = false;
  int pos = -1;
  // the ray waits for the leaf tests as stored (f64 origin, fp32 direction and time): fewer live registers
  double rox = 0., roy = 0., roz = 0.;
  float rdx = 0.f, rdy = 0.f, rdz = 0.f, rtime = 0.f;
  double ddx = 0., ddy = 0., ddz = 0.;  // exact arm only: the f64 direction of a primary ray
  SlabRay sr = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float tbest32 = 0.f;
  const float tmin_lo = __double2float_rd(0.0001), tmin_hi = __double2float_ru(0.0001);
  Hit best;
  hit_reset(best);
  Cands cd;
  cands_reset(cd);
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
#if WF_PREFETCH_AHEAD > 0
        // the queue is streamed front to back by all warps together: ask the L2 for the lines the cursor reaches soon
        // (3 planes x up to 4 lines of 128 B cover the <= 32 positions of one fetch)
        if (lane < 12) {
          const int kp = base + WF_PREFETCH_AHEAD + (lane & 3) * 8;
          if (kp < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(ray_plane(rays_in, Q.capacity, lane >> 2) + kp));
        }
#endif
        if (!have && !exhausted) {
          const int k = base + __popc(need & ((1u << lane) - 1u));
          if (k < n) {
            pos = k;
            const uint4 a = ld_stream(ray_plane(rays_in, Q.capacity, 0) + k), b = ld_stream(ray_plane(rays_in, Q.capacity, 1) + k),
                        c = ld_stream(ray_plane(rays_in, Q.capacity, 2) + k);
            rox = bits_to_double(a.x, a.y); roy = bits_to_double(a.z, a.w); roz = bits_to_double(b.x, b.y);
            if (k >= n_surv) {  // primary record: f64 direction, time in plane 3
              const double x = bits_to_double(b.z, b.w), y = bits_to_double(c.x, c.y), z = bits_to_double(c.z, c.w);
              rdx = (float)x; rdy = (float)y; rdz = (float)z;
              if (!CAND) { ddx = x; ddy = y; ddz = z; }
              rtime = __uint_as_float(__ldcs(reinterpret_cast<const unsigned*>(ray_plane(rays_in, Q.capacity, 3) + k)));
            } else {
              rdx = __uint_as_float(b.z); rdy = __uint_as_float(b.w); rdz = __uint_as_float(c.x); rtime = __uint_as_float(c.y);
              if (!CAND) { ddx = (double)rdx; ddy = (double)rdy; ddz = (double)rdz; }
            }
            sr = NODES == NODES_Q ? slab_ray_q(S, rox, roy, roz, rdx, rdy, rdz) : slab_ray(rox, roy, roz, rdx, rdy, rdz);
            if (CAND) { cands_reset(cd); tbest32 = cd.bound; }
            else { hit_reset(best); tbest32 = __double2float_ru(best.t); }
#if WF_DENSE_LEAVES
            if (CAND) {
              LeafRay lr;
              lr.ox = rox; lr.oy = roy; lr.oz = roz; lr.dx = rdx; lr.dy = rdy; lr.dz = rdz; lr.time = rtime; lr.pad = 0.f;
              leaf_rays[threadIdx.x] = lr;
            }
#endif
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
        const F8 A = load_f8(N + 0), B = load_f8(N + 2), C = load_f8(N + 4), D = load_f8(N + 6);
        const float4 lx = A.lo, hx = A.hi, ly = B.lo, hy = B.hi, lz = C.lo, hz = C.hi;
        const int4 rf = make_int4(__float_as_int(D.lo.x), __float_as_int(D.lo.y), __float_as_int(D.lo.z), __float_as_int(D.lo.w));
        float t0, t1, t2, t3;
        bool h0, h1, h2, h3;
        slab_box(lx.x, hx.x, ly.x, hy.x, lz.x, hz.x, sr, tmin_lo, tbest32, t0, h0);
        slab_box(lx.y, hx.y, ly.y, hy.y, lz.y, hz.y, sr, tmin_lo, tbest32, t1, h1);
        slab_box(lx.z, hx.z, ly.z, hy.z, lz.z, hz.z, sr, tmin_lo, tbest32, t2, h2);
        slab_box(lx.w, hx.w, ly.w, hy.w, lz.w, hz.w, sr, tmin_lo, tbest32, t3, h3);
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
          slab_box_q(q0.x, q0.y, q0.z, sr, tmin_lo, tbest32, tn0, h0);
          slab_box_q(q1.x, q1.y, q1.z, sr, tmin_lo, tbest32, tn1, h1);
          ch0 = (int)q0.w; ch1 = (int)q1.w;
        } else {
          float4 n0, n1, n2, n3;
          if (NODES == NODES_BVH2_SMEM && node < n_top) {
            const float4* N = top_nodes + 4 * node;
            n0 = N[0]; n1 = N[1]; n2 = N[2]; n3 = N[3];
          } else {
            const float4* N = S.nodes + 4 * (size_t)node;
            const F8 A = load_f8(N + 0), B = load_f8(N + 2);
            n0 = A.lo; n1 = A.hi; n2 = B.lo; n3 = B.hi;
          }
          slab_box(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, sr, tmin_lo, tbest32, tn0, h0);
          slab_box(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, sr, tmin_lo, tbest32, tn1, h1);
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
#if WF_BREAK_RELATIVE
      if (5 * __popc(looping) < 2 * entered) break;  // fewer than 40 % of the lanes that entered are still descending
#else
      if (entered - __popc(looping) >= WF_BREAK_LEFT) break;
#endif
#endif
    }
    // ---- leaves: the postponed one, then the current node if it is a leaf too -------------------------
    if (CAND) {
      const PfRay pr = pf_ray(rox, roy, roz, rdx, rdy, rdz, rtime, S.scene_mag);
#if WF_DENSE_LEAVES
      // Dense leaf phase: the leaf references of the whole warp (both slots of every lane, box leaves first) go to a
      // per-warp list in shared memory, and lane j classifies entry j for whichever lane owns it -- the ray comes from
      // the owner's shared-memory record, the result goes back through the list.  The leaf code then runs once per ~32
      // entries at ~20 lanes instead of once per slot and kind at 6-8.
      (void)pr;
      __syncwarp();
      for (;;) {
        const int L0 = leaf < 0 ? leaf : 0, L1 = node < 0 ? node : 0;
        const bool b0 = L0 < 0 && leaf_kind_bits(L0) == LEAF_KIND_BOX, b1 = L1 < 0 && leaf_kind_bits(L1) == LEAF_KIND_BOX;
        const bool o0 = L0 < 0 && !b0, o1 = L1 < 0 && !b1;
        const unsigned mb0 = __ballot_sync(FULL, b0), mb1 = __ballot_sync(FULL, b1), mo0 = __ballot_sync(FULL, o0), mo1 = __ballot_sync(FULL, o1);
        if ((mb0 | mb1 | mo0 | mo1) == 0u) break;
        const int nb0 = __popc(mb0), nb = nb0 + __popc(mb1), no0 = __popc(mo0), E = nb + no0 + __popc(mo1);
        const unsigned lt = (1u << lane) - 1u;
        const int p0 = b0 ? __popc(mb0 & lt) : (o0 ? nb + __popc(mo0 & lt) : -1);
        const int p1 = b1 ? nb0 + __popc(mb1 & lt) : (o1 ? nb + no0 + __popc(mo1 & lt) : -1);
        int4* ent = leaf_list[threadIdx.x >> 5];
        if (p0 >= 0) ent[p0] = make_int4(L0, lane, __float_as_int(cd.bound), 0);
        if (p1 >= 0) ent[p1] = make_int4(L1, lane, __float_as_int(cd.bound), 0);
        __syncwarp();
        for (int base = 0; base < E; base += 32) {
          const int e = base + lane;
          if (e < E) {
            const int4 en = ent[e];
            const LeafRay lr = leaf_rays[(threadIdx.x & ~31) + en.y];
            const PfRay wr = pf_ray(lr.ox, lr.oy, lr.oz, lr.dx, lr.dy, lr.dz, lr.time, S.scene_mag);
            int ref = 0;
            float t_lo = 0.f, t_hi = 0.f;
            const int cls = prefilter_classify<NODES == NODES_BVH2_MULTI>(S, en.x, wr, tmin_lo, tmin_hi, __int_as_float(en.z), ref, t_lo, t_hi);
            ent[e] = make_int4(cls, ref, __float_as_int(t_lo), __float_as_int(t_hi));
          }
        }
        __syncwarp();
        if (p0 >= 0) {
          const int4 r0 = ent[p0];
          cands_apply(cd, r0.x, r0.y, __int_as_float(r0.z), __int_as_float(r0.w));
          if (STATS) st_prims += (unsigned long long)(b0 ? 1 : leaf_count(L0));
          leaf = 0;
        }
        if (p1 >= 0) {
          const int4 r1 = ent[p1];
          cands_apply(cd, r1.x, r1.y, __int_as_float(r1.z), __int_as_float(r1.w));
          if (STATS) st_prims += (unsigned long long)(b1 ? 1 : leaf_count(L1));
          WF_POP();
        }
        if (cd.c0 == CAND_OVERFLOW) { leaf = 0; node = TRAV_DONE; }  // the exact kernel starts over
        __syncwarp();  // the list is rewritten by the next round
      }
#else
      while (leaf < 0) {
        const int count = prefilter_leaf<NODES == NODES_BVH2_MULTI>(S, leaf, pr, tmin_lo, tmin_hi, cd);
        if (STATS) st_prims += (unsigned long long)count;
        leaf = 0;
        if (node < 0) {
          leaf = node;
          WF_POP();
        }
      }
#endif
      if (have) {
        tbest32 = cd.bound;
        if (cd.c0 == CAND_OVERFLOW) node = TRAV_DONE;  // the exact kernel starts over: nothing left to find here
      }
    } else {
      Ray r;
      r.ox = rox; r.oy = roy; r.oz = roz;
      r.dx = ddx; r.dy = ddy; r.dz = ddz; r.time = (double)rtime;
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
    }
#if WF_SYNC_AFTER_LEAVES
    // The lanes leave the leaf phase in groups (no leaf / quad / sphere / second leaf).  Without an explicit
    // reconvergence point ptxas let each group run on to the loop top on its own (ncu: the block below and the
    // fetch prologue executed 3.8 M times at 9 lanes instead of 1.3 M times at 32, and every __ballot_sync went
    // through an out-of-line WARPSYNC.COLLECTIVE path).
    __syncwarp();
#endif
    if (have && node == TRAV_DONE) {
      int2 h;
      if (CAND) {
        cands_record(cd, h.x, h.y);
        if (cd.c0 == CAND_OVERFLOW) Q.overflow[atomicAdd(&Q.c->n_overflow, 1)] = pos;
      } else {  // the winner as the only candidate: the shade stage repeats its test (bit-identical) for t
        h.x = 0; h.y = 0;
        if (best.prim >= 0) {
          const int info_x = __ldg(&S.prim_info[best.prim].x);
          h.x = leaf_make(best.prim, 1, ((info_x & 0xFF) == PRIM_QUAD ? LEAF_KIND_QUAD : 0) | ((info_x & PRIM_FLAG_MOVING) ? LEAF_KIND_MOVING : 0));
        }
      }
      st_stream(Q.cands + pos, h);
      have = false;
    }
  }
  if (STATS) {
    atomicAdd(&stats->node_visits, st_nodes);
    atomicAdd(&stats->prim_tests, st_prims);
  }
}

// the ray at queue position `pos`, whichever its encoding
__device__ __forceinline__ void load_path(const RayRec* __restrict__ rays, int cap, int pos, PathRec& p) {
  const uint4 a = __ldg(ray_plane(rays, cap, 0) + pos), b = __ldg(ray_plane(rays, cap, 1) + pos);
  const uint4 c = __ldg(ray_plane(rays, cap, 2) + pos), d = __ldg(ray_plane(rays, cap, 3) + pos);
  unpack(a, b, c, d, p);
}

// Rays whose live candidates did not fit the slots (coplanar faces, grazing hits: ~0.05 % on c4): closest hit by the
// exact traversal (closest_surface: fp32 cull, f64 leaves), the winner becomes the only candidate.
// (67 us of mostly latency per iteration, 1.9 % of the step.  Folding the re-trace into k_wf_shade_rare -- so that it
//  overlaps the deferred items -- was measured: 440.7 vs 438.3 ms per c4 step, no gain, and the re-traced rays would be
//  shaded by another instantiation than in the exact-leaves arm, which costs the bit-identity of the two arms.)
__global__ void __launch_bounds__(128) k_wf_extend_exact(const __grid_constant__ DScene S, WFQueues Q, const RayRec* __restrict__ rays_in) {
  const int n = Q.c->n_overflow;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int pos = Q.overflow[i];
    PathRec p;
    load_path(rays_in, Q.capacity, pos, p);
    Hit best;
    hit_reset(best);
    closest_surface<false>(S, to_ray(p), 0.0001, best, nullptr);
    int2 h = make_int2(0, 0);
    if (best.prim >= 0) {
      const int info_x = __ldg(&S.prim_info[best.prim].x);
      h.x = leaf_make(best.prim, 1, ((info_x & 0xFF) == PRIM_QUAD ? LEAF_KIND_QUAD : 0) | ((info_x & PRIM_FLAG_MOVING) ? LEAF_KIND_MOVING : 0));
    }
    Q.cands[pos] = h;
  }
}

// ------------------------------------------------------------------------------------------------
// shade (fused resolve + sort + shade).  One block owns WF_SHADE_BLOCK consecutive queue positions:
//   1. stream the ray + candidate records in (coalesced), run the exact reference-order tests on the
//      candidates, add the constant-medium events (ConstantMedium::hit, constant_medium.rs:41-95) and look
//      up the shading class;
//   2. block-local counting sort by class through shared memory, so that a warp shades one class
//      (the Perlin texture costs ~1k instructions, a solid Lambertian ~150: they must not share a warp);
//   3. shade (ray_color's match arms, render.rs:271-297); survivors are appended densely to the next
//      queue with one atomic per warp, finished paths add their radiance to the image.
// No global bins, no gathers of queue records: every queue access of this kernel is sequential.
// ------------------------------------------------------------------------------------------------
constexpr int WF_SHADE_BLOCK = 128;  // measured on c4: 128 -> 23.0, 256 -> 23.9, 512 -> 26.0 ms shade per step
#ifndef WF_SHADE_MIN_BLOCKS
#define WF_SHADE_MIN_BLOCKS 7
#endif

struct ShadeItem {  // what moves through shared memory to the lane that shades it (80 B)
  uint4 a, b, c, d;
  double t;
  int id;
  int info_x;  // prim_info[id].x of a surface hit (kind | flags | class | material)
};

// 1. closest surface hit from the candidates, then the constant-medium events (each lane for its own ray), then
//    the shading class.  The f64 boundary intervals run at ~13 of 32 lanes (only the rays whose fp32
//    rejections do not settle it); queueing those (ray, medium) pairs in shared memory and evaluating the
//    queue densely after a barrier was measured: 18.8 -> 23.1 ms per c4 row.  Kept per-lane.
template <bool STATS, int SPEC>
__device__ __forceinline__ int wf_resolve(const DScene& S, const Tables& T, const uint4& a, const uint4& b, const uint4& c,
                                          const uint4& d, int2 cand, double& t, int& id, int& info_x, DStats& st) {
  t = RTB_INF; id = -1; info_x = 0;
  if (d.y == PADDING_PIXEL) return CLS_MISS;
  PathRec p;
  unpack(a, b, c, d, p);
  const Ray r = to_ray(p);
  if (cand.x < 0) {
    Hit best;
    resolve_candidates<(SPEC & SPEC_MULTI_LEAF) != 0>(S, cand.x, cand.y, r, 0.0001, best);
    if (STATS) st.exact_tests += (unsigned long long)((cand.x < 0 ? leaf_count(cand.x) : 0) + (cand.y < 0 ? leaf_count(cand.y) : 0));
    if (best.prim >= 0) {
      t = best.t; id = best.prim;
      info_x = RTB_LDG(&S.prim_info[best.prim].x);
    }
  }
  if ((SPEC & SPEC_MEDIA) && S.n_media > 0) {
    Rand4 u;
    for (int mi = 0; mi < S.n_media; mi++) {
      if ((mi & 3) == 0) u = rand4(S, p.pixel, p.sample, p.bounce, 1u + (uint32_t)(mi >> 2));
      const float U = (mi & 3) == 0 ? u.x : ((mi & 3) == 1 ? u.y : ((mi & 3) == 2 ? u.z : u.w));
      const double tm = medium_event_lazy<(SPEC & SPEC_BOXSCAN) != 0, (SPEC & SPEC_GENERIC_MEDIA) != 0>(S, T.media[mi], r, 0.0001, t, U);
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

// one finished path: 64-bit fixed-point adds (order-independent -> bit-reproducible sums)
__device__ __forceinline__ void wf_accumulate(const DScene& S, unsigned long long* __restrict__ accum, uint32_t pixel, float Lr, float Lg, float Lb,
                                              DStats& st, bool stats) {
  const bool finite = (fabsf(Lr) < 3.0e38f) && (fabsf(Lg) < 3.0e38f) && (fabsf(Lb) < 3.0e38f);
  unsigned long long* acc = accum + 4ull * pixel;
  if (finite) {
    if (Lr != 0.f) atomicAdd(acc + 0, accum_fixed(Lr));
    if (Lg != 0.f) atomicAdd(acc + 1, accum_fixed(Lg));
    if (Lb != 0.f) atomicAdd(acc + 2, accum_fixed(Lb));
  } else if (S.flags & 2u) {
    atomicOr(acc + 3, ACCUM_POISON);  // HEAD-literal (Q22): a non-finite sample poisons the pixel
  } else if (stats) {
    st.nonfinite++;
  }
}

// 3. ray_color's match arms for one item (render.rs:271-297): finished paths add their radiance to the image,
//    survivors return true with the next ray packed into `out`.
template <bool STATS, int SPEC>
__device__ __forceinline__ bool wf_shade_item(const DScene& S, const Tables& T, const uint4& a, const uint4& b, const uint4& c,
                                              const uint4& d, double t, int id, int info_x, unsigned long long* __restrict__ accum,
                                              RayRec& out, DStats& st) {
  if (d.y == PADDING_PIXEL) return false;
  PathRec p;
  unpack(a, b, c, d, p);
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
    p.dx = ps.ray.dx; p.dy = ps.ray.dy; p.dz = ps.ray.dz;
    p.bx = ps.bx; p.by = ps.by; p.bz = ps.bz;
    p.bounce = ps.bounce;
    out = pack_secondary(p);
    return true;
  }
  wf_accumulate(S, accum, p.pixel, Lr, Lg, Lb, st, STATS);
  // (the per-pixel sample count, accum[3], is added in bulk by k_wf_add_count: every pixel receives
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

__device__ __forceinline__ void flush_shade_stats(const DStats& st, DStats* stats) {
  if (st.nonfinite) atomicAdd(&stats->nonfinite, st.nonfinite);
  if (st.exact_tests) atomicAdd(&stats->exact_tests, st.exact_tests);
}

// Scene-specialised instantiations (chosen at launch from DScene::spec_bits, device_scene.h SPEC_*): code a scene
// never runs still costs it registers and instruction-cache misses in this 70 KB kernel (c4: +3.7 %).
// DEFER: items of the rare, heavy, textured classes (image / Perlin Lambertians: 0.8 % of c4's items, but 8 % of this
// kernel's warp instructions at 5 of 32 lanes, and 20 KB of its code) are not shaded here: their queue positions
// go to a list that k_wf_shade_rare works off densely, and this instantiation carries no texture code at all.
template <bool STATS, int SPEC, bool DEFER = false>
__global__ void __launch_bounds__(WF_SHADE_BLOCK, (SPEC & SPEC_GENERIC_MEDIA) ? WF_SHADE_MIN_BLOCKS : WF_SHADE_MIN_BLOCKS + 1) k_wf_shade(const __grid_constant__ DScene S, WFQueues Q,
                                                            const RayRec* __restrict__ rays_in,
                                                            RayRec* __restrict__ rays_out, unsigned long long* __restrict__ accum,
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
  DStats st = {};
  // ---- 1. load + exact candidate tests + medium events + class ------------------------------------------
  ShadeItem it;
  int cls = -1;
  const int cap = Q.capacity;
  int2 cand = make_int2(0, 0);
  if (i < n) {
    it.a = ld_stream(ray_plane(rays_in, cap, 0) + i); it.b = ld_stream(ray_plane(rays_in, cap, 1) + i);
    it.c = ld_stream(ray_plane(rays_in, cap, 2) + i); it.d = ld_stream(ray_plane(rays_in, cap, 3) + i);
    cand = ld_stream(Q.cands + i);
  }
  __syncthreads();  // counters zeroed
  if (i < n) cls = wf_resolve<STATS, SPEC>(S, T, it.a, it.b, it.c, it.d, cand, it.t, it.id, it.info_x, st);
  if (DEFER) {
    // a deferred item is a surface hit no medium event came before: its resolved hit goes to hits[i]
    const bool defer = (cls == CLS_LAMBERT_TEX || cls == CLS_NOISE) && it.id >= 0;
    const unsigned dm = __ballot_sync(0xFFFFFFFFu, defer);
    if (dm) {
      const int leader = __ffs(dm) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&Q.c->n_deferred, __popc(dm));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (defer) {
        Q.deferred[base + __popc(dm & ((1u << lane) - 1u))] = i;
        HitRec h;
        h.t = it.t; h.id = it.id; h.info_x = it.info_x;
        Q.hits[i] = h;
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
  if (tid < n_sorted) {
    const ShadeItem me = items[tid];
    alive = wf_shade_item<STATS, SPEC>(S, T, me.a, me.b, me.c, me.d, me.t, me.id, me.info_x, accum, out, st);
  }
  wf_append(Q, rays_out, alive, out, lane);
  if (STATS) {
    flush_shade_stats(st, stats);
    if (tid == 0 && S.n_media > 0) atomicAdd(&stats->medium_probes, (unsigned long long)S.n_media * (unsigned long long)n_block);
  }
}

// the deferred items of k_wf_shade<.., DEFER = true>: same shading (every feature compiled in), dense lanes
template <bool STATS>
__global__ void __launch_bounds__(WF_SHADE_BLOCK) k_wf_shade_rare(const __grid_constant__ DScene S, WFQueues Q,
                                                                   const RayRec* __restrict__ rays_in, RayRec* __restrict__ rays_out,
                                                                   unsigned long long* __restrict__ accum, DStats* __restrict__ stats) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int nd = Q.c->n_deferred, cap = Q.capacity;
  const Tables T = scene_tables(S);
  DStats st = {};
  for (int base = blockIdx.x * WF_SHADE_BLOCK; base < nd; base += gridDim.x * WF_SHADE_BLOCK) {  // uniform per block
    bool alive = false;
    RayRec out;
    if (base + tid < nd) {
      const int j = Q.deferred[base + tid];
      const uint4 a = __ldg(ray_plane(rays_in, cap, 0) + j), b = __ldg(ray_plane(rays_in, cap, 1) + j);
      const uint4 c = __ldg(ray_plane(rays_in, cap, 2) + j), d = __ldg(ray_plane(rays_in, cap, 3) + j);
      const uint4 h = *reinterpret_cast<const uint4*>(Q.hits + j);
      alive = wf_shade_item<STATS, SPEC_ALL>(S, T, a, b, c, d, bits_to_double(h.x, h.y), (int)h.z, (int)h.w, accum, out, st);
    }
    wf_append(Q, rays_out, alive, out, lane);
  }
  if (STATS) flush_shade_stats(st, stats);
}

// ------------------------------------------------------------------------------------------------
// finish: the decaying tail of a call.  Once every path has started and few are left, each further
// iteration still costs its ~50 us of launches and kernel tails (a depth-50 scene pays 40 of them for a
// handful of rays).  This kernel runs each remaining path to its end in one launch -- the megakernel's
// loop (exact traversal + shade per lane) over the queue entries.  Same segments, same sums.
// ------------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) k_wf_finish(const __grid_constant__ DScene S, WFQueues Q, const RayRec* __restrict__ rays_in,
                                                   unsigned long long* __restrict__ accum, DStats* __restrict__ stats) {
  const int n = Q.c->n_out;  // the queue the last shade + generate left (n_in of the iteration that will not run)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PathRec p;
  load_path(rays_in, Q.capacity, i, p);
  if (p.pixel == PADDING_PIXEL) return;
  PathState ps;
  ps.ray = to_ray(p);
  ps.bx = p.bx; ps.by = p.by; ps.bz = p.bz;
  ps.pixel = p.pixel; ps.sample = p.sample; ps.bounce = p.bounce;
  DStats st = {};
  float Lr = 0.f, Lg = 0.f, Lb = 0.f;
  bool alive = true;
  RTB_LOOP_ENTER();
  while (alive) {
    Event ev;
    if (STATS) st.segments++;
    extend<STATS>(S, ps, ev, &st);
    alive = shade(S, ps, ev, Lr, Lg, Lb, &st, STATS);  // (scattered directions are fp32-valued, as in the queues)
  }
  RTB_LOOP_LEAVE();
  wf_accumulate(S, accum, p.pixel, Lr, Lg, Lb, st, STATS);
  if (STATS) {
    atomicAdd(&stats->segments, st.segments);
    atomicAdd(&stats->node_visits, st.node_visits);
    atomicAdd(&stats->prim_tests, st.prim_tests);
    atomicAdd(&stats->medium_probes, st.medium_probes);
    if (st.nonfinite) atomicAdd(&stats->nonfinite, st.nonfinite);
  }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
template <int SPEC>
static void launch_shade_spec(const DScene& S, const WFQueues& Q, const RayRec* in, RayRec* out, unsigned long long* d_accum, DStats* d_stats,
                              unsigned blocks, cudaStream_t st) {
  k_wf_shade<false, SPEC><<<blocks, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
}
template <int SPEC>
static void launch_shade_defer(const DScene& S, const WFQueues& Q, const RayRec* in, RayRec* out, unsigned long long* d_accum, DStats* d_stats,
                               unsigned blocks, int sms, cudaStream_t st) {
  k_wf_shade<false, SPEC, true><<<blocks, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
  const unsigned rare = std::max(1u, std::min(blocks / 16u, (unsigned)(sms * 4)));
  k_wf_shade_rare<false><<<rare, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
}
// returns the number of kernels launched
static int launch_shade(const DScene& S, const WFQueues& Q, const RayRec* in, RayRec* out, unsigned long long* d_accum, DStats* d_stats,
                        bool collect_stats, unsigned blocks, bool defer_rare, int sms, cudaStream_t st) {
  constexpr int kTex = SPEC_TEXTURES | SPEC_SPHERE_UV | SPEC_QUAD_UV;
  const int need = S.spec_bits & SPEC_ALL;
  if (collect_stats || (need & SPEC_MULTI_LEAF)) {  // counted passes are not timed, multi-primitive leaves are a tuning arm: one generic instantiation
    if (collect_stats) k_wf_shade<true, SPEC_ALL><<<blocks, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
    else k_wf_shade<false, SPEC_ALL><<<blocks, WF_SHADE_BLOCK, 0, st>>>(S, Q, in, out, d_accum, d_stats);
    return 1;
  }
  if (defer_rare && S.defer_ok && (need & SPEC_TEXTURES)) {
    const int rest = need & ~kTex;
    if (rest == SPEC_MEDIA) { launch_shade_defer<SPEC_MEDIA>(S, Q, in, out, d_accum, d_stats, blocks, sms, st); return 2; }
    if (rest == (SPEC_MEDIA | SPEC_LIGHTS)) { launch_shade_defer<SPEC_MEDIA | SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, sms, st); return 2; }
    if (rest == 0) { launch_shade_defer<0>(S, Q, in, out, d_accum, d_stats, blocks, sms, st); return 2; }
  }
  // the smallest instantiation whose features cover the scene's (a superset is always correct)
  constexpr int kC4 = SPEC_MEDIA | SPEC_SPHERE_UV | SPEC_TEXTURES, kC3 = SPEC_MEDIA | SPEC_BOXSCAN | SPEC_GENERIC_MEDIA;
  constexpr int kAll = SPEC_ALL & ~SPEC_MULTI_LEAF;
  auto covers = [need](int spec) { return (spec & need) == need; };
  if (covers(0)) launch_shade_spec<0>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(SPEC_LIGHTS)) launch_shade_spec<SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC4)) launch_shade_spec<kC4>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC3)) launch_shade_spec<kC3>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC4 | SPEC_LIGHTS)) launch_shade_spec<kC4 | SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else if (covers(kC3 | SPEC_LIGHTS)) launch_shade_spec<kC3 | SPEC_LIGHTS>(S, Q, in, out, d_accum, d_stats, blocks, st);
  else launch_shade_spec<kAll>(S, Q, in, out, d_accum, d_stats, blocks, st);
  return 1;
}

template <bool STATS, bool CAND>
static void launch_extend_nodes(const DScene& S, const WFQueues& Q, const RayRec* in, DStats* d_stats, unsigned grid, bool smem_top, cudaStream_t st) {
  if (S.multi_leaf) k_wf_extend<STATS, NODES_BVH2_MULTI, CAND><<<grid, WF_EXTEND_BLOCK, 0, st>>>(S, Q, in, d_stats);
  else if (S.use_qnodes) k_wf_extend<STATS, NODES_Q, CAND><<<grid, WF_EXTEND_BLOCK, 0, st>>>(S, Q, in, d_stats);
  else if (S.use_bvh4) k_wf_extend<STATS, NODES_BVH4, CAND><<<grid, WF_EXTEND_BLOCK, 0, st>>>(S, Q, in, d_stats);
  else if (smem_top) k_wf_extend<STATS, NODES_BVH2_SMEM, CAND><<<grid, WF_EXTEND_BLOCK, 0, st>>>(S, Q, in, d_stats);
  else k_wf_extend<STATS, NODES_BVH2, CAND><<<grid, WF_EXTEND_BLOCK, 0, st>>>(S, Q, in, d_stats);
}
static void launch_extend(const DScene& S, const WFQueues& Q, const RayRec* in, DStats* d_stats, unsigned grid, bool stats, bool cand, bool smem_top,
                          cudaStream_t st) {
  if (stats) { if (cand) launch_extend_nodes<true, true>(S, Q, in, d_stats, grid, smem_top, st); else launch_extend_nodes<true, false>(S, Q, in, d_stats, grid, smem_top, st); }
  else { if (cand) launch_extend_nodes<false, true>(S, Q, in, d_stats, grid, smem_top, st); else launch_extend_nodes<false, false>(S, Q, in, d_stats, grid, smem_top, st); }
}
template <int NODES>
static int extend_occupancy(bool stats, bool cand) {
  int r = 0;
  if (stats) { if (cand) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_wf_extend<true, NODES, true>, WF_EXTEND_BLOCK, 0); else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_wf_extend<true, NODES, false>, WF_EXTEND_BLOCK, 0); }
  else { if (cand) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_wf_extend<false, NODES, true>, WF_EXTEND_BLOCK, 0); else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_wf_extend<false, NODES, false>, WF_EXTEND_BLOCK, 0); }
  return r;
}
static int extend_blocks_per_sm(const DScene& S, bool stats, bool cand, bool smem_top) {
  int r = S.multi_leaf ? extend_occupancy<NODES_BVH2_MULTI>(stats, cand)
          : S.use_qnodes ? extend_occupancy<NODES_Q>(stats, cand)
          : S.use_bvh4 ? extend_occupancy<NODES_BVH4>(stats, cand)
          : smem_top ? extend_occupancy<NODES_BVH2_SMEM>(stats, cand) : extend_occupancy<NODES_BVH2>(stats, cand);
  return r < 1 ? 1 : r;
}

// accum[3] += number of strata rendered, for every pixel (what one atomicAdd(+1) per finished path would sum to)
__global__ void k_wf_add_count(unsigned long long* __restrict__ accum, int n_pixels, unsigned long long n_strata) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pixels) accum[4ull * i + 3] += n_strata;
}

cudaError_t launch_add_count(unsigned long long* d_accum, int64_t n_pixels, unsigned long long n_strata, cudaStream_t stream) {
  if (n_pixels <= 0) return cudaSuccess;
  k_wf_add_count<<<(unsigned)((n_pixels + 255) / 256), 256, 0, stream>>>(d_accum, (int)n_pixels, n_strata);
  return cudaGetLastError();
}

// (the queue entries of the pixels that pad the image to whole 8 x 4 tiles are dropped by their first shade: one entry each,
//  not a path segment)
__global__ void k_wf_sum_counters(WFQueues Q, DStats* stats, unsigned long long padding_entries) {
  stats->segments += Q.c->segments - padding_entries;
  stats->overflows += Q.c->overflows;
}

cudaError_t launch_render_wavefront(const DScene& S, const WavefrontContext& ctx, const WavefrontOptions& opt, int64_t s_begin, int64_t s_end,
                                    unsigned long long* d_accum, DStats* d_stats, bool collect_stats, void* d_workspace,
                                    size_t workspace_bytes, int64_t capacity, cudaStream_t stream, int* launches, double* stage_out) {
  if (workspace_bytes < wavefront_workspace_bytes(S, capacity)) return cudaErrorInvalidValue;
  const long long n_strata = (long long)(s_end - s_begin);
  if (n_strata < 1) return cudaSuccess;
  const bool profile = opt.profile != 0;  // per-stage CUDA-event totals on stderr (analysis runs only)
  const bool cand = opt.exact_leaves == 0;
  const bool smem_top = opt.smem_top != 0;
  const int64_t cap = capacity & ~(int64_t)255;
  const unsigned long long tiles = (unsigned long long)((S.cam.width + 7) / 8) * ((S.cam.height + 3) / 4);
  int per_sm = extend_blocks_per_sm(S, collect_stats, cand, smem_top);
  if (opt.extend_blocks_per_sm > 0) per_sm = std::min(per_sm, opt.extend_blocks_per_sm);
  const long long extend_grid_full = (long long)ctx.sms * per_sm;
  WFQueues Q = carve(d_workspace, cap);
  RayRec* in = Q.rays_a;
  RayRec* out = Q.rays_b;
  cudaError_t e;
  int n_launch = 0;
  k_wf_init<<<1, 1, 0, stream>>>(Q, tiles * 32ull * (unsigned long long)n_strata);
  n_launch++;
  WFCounters* h_c = static_cast<WFCounters*>(ctx.host_counters);  // pinned mirror for the (sparse) host polls
  cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
  double stage_ms[3] = {0., 0., 0.};
  if (profile)
    for (auto& ev : pe) cudaEventCreate(&ev);
  int poll_every = profile ? 1 : 8;
  long long iters = 0;
  // Upper bound of the rays of the coming iterations, known to the host only at polls: while paths are
  // still being started the queue is full; once every path has started it can only shrink, so the last
  // polled count bounds all later iterations and the launches of the decaying tail are sized to it.
  long long bound = cap;
  const long long finish_below = opt.finish_below >= 0 ? opt.finish_below : 65536;
  bool finished = false;
  for (long long iter = 0;; iter++) {
    const unsigned gen_blocks = (unsigned)std::min<long long>((bound + 255) / 256, (long long)ctx.sms * 8);
    const unsigned shade_blocks = (unsigned)((bound + WF_SHADE_BLOCK - 1) / WF_SHADE_BLOCK);
    const unsigned extend_grid = (unsigned)std::min<long long>(extend_grid_full, (bound + WF_EXTEND_BLOCK - 1) / WF_EXTEND_BLOCK);
    // top the out queue up (first iteration: fill it), then it becomes this iteration's in queue
    if (profile) cudaEventRecord(pe[0], stream);
    k_wf_generate<<<gen_blocks, 256, 0, stream>>>(S, Q, s_begin, out);
    if (profile) cudaEventRecord(pe[1], stream);
    { RayRec* t = in; in = out; out = t; }
    launch_extend(S, Q, in, d_stats, extend_grid, collect_stats, cand, smem_top, stream);
    n_launch += 2;
    if (cand) {
      k_wf_extend_exact<<<(unsigned)std::min<long long>(ctx.sms, (bound + 127) / 128), 128, 0, stream>>>(S, Q, in);
      n_launch++;
    }
    if (profile) cudaEventRecord(pe[2], stream);
    n_launch += launch_shade(S, Q, in, out, d_accum, d_stats, collect_stats, shade_blocks, opt.defer_rare != 0, ctx.sms, stream);
    if (profile) {
      cudaEventRecord(pe[3], stream);
      cudaEventSynchronize(pe[3]);
      float it_ms[3];
      for (int j = 0; j < 3; j++) { it_ms[j] = 0.f; cudaEventElapsedTime(&it_ms[j], pe[j], pe[j + 1]); stage_ms[j] += it_ms[j]; }
      if (opt.profile >= 2) fprintf(stderr, "[rtb iter] %lld extend %.3f shade %.3f\n", iter, it_ms[1], it_ms[2]);
    }
    iters++;
    if ((iter % poll_every) == poll_every - 1) {
      if ((e = cudaMemcpyAsync(h_c, Q.c, sizeof(WFCounters), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
      if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
      if (h_c->next_path >= h_c->total_paths) {
        if (h_c->n_out == 0) break;
        bound = std::max<long long>(h_c->n_out, 1);  // every path has started: the queue only shrinks
        if (bound <= finish_below) {  // few paths left: one kernel runs them to their end
          const unsigned blocks = (unsigned)((bound + 127) / 128);
          if (collect_stats) k_wf_finish<true><<<blocks, 128, 0, stream>>>(S, Q, out, d_accum, d_stats);
          else k_wf_finish<false><<<blocks, 128, 0, stream>>>(S, Q, out, d_accum, d_stats);
          n_launch++;
          finished = true;
          break;
        }
        if (!profile) poll_every = bound < cap / 4 ? 2 : 4;
      }
    }
  }
  if (profile) {
    fprintf(stderr, "[rtb wavefront] iterations %lld%s  segments %llu  overflows %llu  generate %.2f ms  extend %.2f ms  shade %.2f ms\n", iters,
            finished ? " + finish" : "", (unsigned long long)h_c->segments, (unsigned long long)h_c->overflows, stage_ms[0], stage_ms[1], stage_ms[2]);
    for (auto& ev : pe) cudaEventDestroy(ev);
    if (stage_out)
      for (int j = 0; j < 3; j++) stage_out[j] = stage_ms[j];
  }
  {
    const int n_pixels = S.cam.width * S.cam.height;
    k_wf_add_count<<<(n_pixels + 255) / 256, 256, 0, stream>>>(d_accum, n_pixels, (unsigned long long)n_strata);
    n_launch++;
  }
  if (collect_stats) {
    const unsigned long long padding = (tiles * 32ull - (unsigned long long)S.cam.width * (unsigned long long)S.cam.height) * (unsigned long long)n_strata;
    k_wf_sum_counters<<<1, 1, 0, stream>>>(Q, d_stats, padding);
    n_launch++;
  }
  if (launches) *launches += n_launch;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// parity harness THROUGH the pipeline's own kernels (rtb_trace with RTB_TRACE_WAVEFRONT): caller rays are
// packed into a real queue (primary records = f64 directions, or secondary records = directions rounded to
// fp32), traversed by k_wf_extend, re-traced by k_wf_extend_exact where the candidates overflowed, resolved by
// the exact tests as the shade stage does, and completed to full hit records.
// ------------------------------------------------------------------------------------------------
__global__ void k_wf_trace_pack(const RtbRay* __restrict__ rays, int n, int secondary, WFQueues Q, RayRec* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    WFCounters z = {};
    z.n_in = n;
    z.n_surv = secondary ? n : 0;
    *Q.c = z;
  }
  if (i >= n) return;
  const RtbRay in = rays[i];
  PathRec p;
  p.ox = in.origin[0]; p.oy = in.origin[1]; p.oz = in.origin[2];
  p.dx = in.direction[0]; p.dy = in.direction[1]; p.dz = in.direction[2];
  p.time = (float)in.time;
  p.bx = p.by = p.bz = 1.f;
  p.pixel = (uint32_t)i; p.sample = 0u; p.bounce = secondary ? 1u : 0u;
  const RayRec rec = secondary ? pack_secondary(p) : pack_primary(p);
  const int cap = Q.capacity;
  ray_plane(out, cap, 0)[i] = rec.a; ray_plane(out, cap, 1)[i] = rec.b;
  ray_plane(out, cap, 2)[i] = rec.c; ray_plane(out, cap, 3)[i] = rec.d;
}

__global__ void k_wf_trace_resolve(const __grid_constant__ DScene S, WFQueues Q, const RayRec* __restrict__ rays_in, int n,
                                   RtbHit* __restrict__ hits, int* __restrict__ cand_counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PathRec p;
  load_path(rays_in, Q.capacity, i, p);
  const Ray r = to_ray(p);
  const int2 cand = Q.cands[i];
  Hit best;
  resolve_candidates<true>(S, cand.x, cand.y, r, 0.0001, best);
  if (cand.y == CAND_CERTAIN && best.prim >= 0 && (leaf_kind_bits(cand.x) & LEAF_KIND_QUAD)) {
    double tt;  // the full record wants (u, v) = (alpha, beta): the shade stage evaluates them only for uv-textured quads
    quad_test(S.prims + (size_t)best.prim * PRIM_D2, r, -RTB_INF, RTB_INF, tt, best.a, best.b);
  }
  RtbHit out;
  complete_hit(S, r, best, out);
  hits[i] = out;
  if (cand_counts) cand_counts[i] = (cand.x < 0) + (cand.y < 0);
}

cudaError_t launch_trace_wavefront(const DScene& S, const WavefrontContext& ctx, const WavefrontOptions& opt, const RtbRay* d_rays, int64_t n,
                                   bool secondary, RtbHit* d_hits, int* d_cand_counts, void* d_workspace, size_t workspace_bytes,
                                   unsigned long long* overflows, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int64_t cap = (n + 255) & ~(int64_t)255;
  if (workspace_bytes < wavefront_workspace_bytes(S, cap)) return cudaErrorInvalidValue;
  WFQueues Q = carve(d_workspace, cap);
  const bool cand = opt.exact_leaves == 0, smem_top = opt.smem_top != 0;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  k_wf_trace_pack<<<blocks, 128, 0, stream>>>(d_rays, (int)n, secondary ? 1 : 0, Q, Q.rays_a);
  if (S.n_surface_prims > 0) {
    const long long grid_full = (long long)ctx.sms * extend_blocks_per_sm(S, false, cand, smem_top);
    launch_extend(S, Q, Q.rays_a, nullptr, (unsigned)std::min<long long>(grid_full, blocks), false, cand, smem_top, stream);
    if (cand) k_wf_extend_exact<<<(unsigned)std::min<long long>(ctx.sms, blocks), 128, 0, stream>>>(S, Q, Q.rays_a);
  } else {
    cudaMemsetAsync(Q.cands, 0, (size_t)n * sizeof(int2), stream);
  }
  k_wf_trace_resolve<<<blocks, 128, 0, stream>>>(S, Q, Q.rays_a, (int)n, d_hits, d_cand_counts);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (overflows) {
    WFCounters* h_c = static_cast<WFCounters*>(ctx.host_counters);
    if ((e = cudaMemcpyAsync(h_c, Q.c, sizeof(WFCounters), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    *overflows = (unsigned long long)h_c->n_overflow;
  }
  return cudaSuccess;
}

}  // namespace rtb
