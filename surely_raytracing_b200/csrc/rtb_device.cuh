// rtb_device.cuh -- device-side building blocks of the path-tracing hot path: Philox RNG,
// f64 primitive tests, fp32 conservative BVH traversal, media, textures, pdfs, materials.
//
// Behavioural contract = reference src/render.rs:251-312 (ray_color) and everything it calls;
// each function cites what it restates.  The arithmetic is this backend's own design:
//   * instance transforms are baked, so there is one world-space BVH2 (no per-instance ray xform);
//   * BVH slabs are fp32 over outward-padded boxes: a conservative cull only;
//   * primitive tests / medium intervals / light-pdf probes are f64 with the reference's interval
//     rules (quads closed, spheres open: Q4) -- B200 issues DFMA at half the FFMA rate, which makes
//     "decide in f64" affordable and is what holds first-hit ids exact against the f64 reference;
//   * shading (ONB, sampling maps, optics, textures) is fp32;
//   * RNG is counter-based Philox4x32-10 keyed (pixel, sample, bounce) with fixed draw slots.
//
// The header also compiles as plain C++ (tests/emu) so the logic can be debugged without a GPU;
// that build is a development aid and is never linked into librtb200.so.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/rtb200.h"
#include "device_scene.h"

#if defined(RTB_DEBUG_BOUNDS)
#include <assert.h>
#define RTB_ASSERT(c) assert(c)
#else
#define RTB_ASSERT(c) ((void)0)
#endif

#if defined(__CUDACC__)
#define RTB_DEV __device__ __forceinline__
// (Turning the rarely-taken paths -- generic medium boundaries, Perlin turbulence, light sampling, sphere uv --
// into real calls to shrink the 73 KB shade kernel was measured: 20.1 -> 26.7 ms per c4 row.  Everything inlines.)
#define RTB_LDG(p) __ldg(p)
#else
#include <cmath>
#include <cstring>
#define RTB_DEV inline
#define RTB_LDG(p) (*(p))
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __double2float_ru(double d) { float f = (float)d; return ((double)f < d) ? nextafterf(f, INFINITY) : f; }
static inline float __double2float_rd(double d) { float f = (float)d; return ((double)f > d) ? nextafterf(f, -INFINITY) : f; }
static inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979323846f * x); *c = cosf(3.14159265358979323846f * x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {  // PRMT, default mode
  const uint64_t src = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((src >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
#endif

namespace rtb {

// Lanes leave a divergent loop one by one.  ptxas keeps loop invariants in UNIFORM registers -- one copy per warp --
// and lets the early leavers run on into code that reuses those registers (BREAK + a branch past the BSYNC): in round 1
// that sent the lanes still traversing to a wild address (DESIGN.md "The -O3 fault"; tools/sass_lint.py).  Every
// divergent loop of the path therefore ends in an explicit reconvergence of the lanes that entered it together.
#if defined(__CUDACC__)
#define RTB_LOOP_ENTER() const unsigned rtb_loop_mask__ = __activemask()
#define RTB_LOOP_LEAVE() __syncwarp(rtb_loop_mask__)
#else
#define RTB_LOOP_ENTER() ((void)0)
#define RTB_LOOP_LEAVE() ((void)0)
#endif

constexpr float PI_F = 3.14159265358979323846f;
constexpr double PI_D = 3.14159265358979323846;
constexpr double RTB_INF = __builtin_huge_val();  // +inf
constexpr uint32_t PRIMARY_BOUNCE = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (pixel, sample, bounce, call), key = seed.
// Replaces rand's ThreadRng behind src/utils.rs:5-15 (SURVEY Appendix A for the slot budget).
// ------------------------------------------------------------------------------------------------
struct Rand4 { float x, y, z, w; };

RTB_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
RTB_DEV float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }  // 24 bits in [0,1)
RTB_DEV Rand4 rand4(const DScene& S, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t call) {
  uint32_t o[4];
  philox4x32_10(pixel, sample, bounce, call, S.seed_lo, S.seed_hi, o);
  Rand4 r;
  r.x = u01(o[0]); r.y = u01(o[1]); r.z = u01(o[2]); r.w = u01(o[3]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// small vector helpers (fp32 shading math)
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
RTB_DEV V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RTB_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RTB_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RTB_DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
RTB_DEV V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RTB_DEV float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
RTB_DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RTB_DEV V3 normalize(V3 a) { return rsqrtf(dot(a, a)) * a; }

// ------------------------------------------------------------------------------------------------
// f64 arithmetic WITHOUT fused multiply-add, in the reference's operation order.  The reference is
// compiled Rust: `a*b + c` is two roundings.  Symmetric scenes put primary rays exactly on shared
// quad edges (the Cornell camera sends its diagonal pixels through the wall/floor corner), where the
// inside test `a < 0` is decided by the last bit -- so the primitive tests mirror the reference's
// arithmetic operation by operation (nvcc would otherwise contract to DFMA).
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
RTB_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
RTB_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
RTB_DEV double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
RTB_DEV double dmul(double a, double b) { return a * b; }   // host build: -ffp-contract=off
RTB_DEV double dadd(double a, double b) { return a + b; }
RTB_DEV double dsub(double a, double b) { return a - b; }
#endif
// dot(u, v) = u.x*v.x + u.y*v.y + u.z*v.z, left to right  (src/vec3.rs:167-169)
RTB_DEV double ddot(double ax, double ay, double az, double bx, double by, double bz) {
  return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
}

struct Ray {
  double ox, oy, oz;  // origin: f64 (hit points must stay on their surface to ~1e-13, like the reference)
  double dx, dy, dz;  // direction, not normalised (Q3); secondary directions are fp32-valued
  double time;        // f64 so that the parity harness sees the reference's centre(time) exactly
};

// The small tagged tables of a scene.  The wavefront shade kernel stages them in shared memory (a
// dependent chain of global gathers per item otherwise); every other caller points at the scene's own.
struct Tables {
  const DMaterial* materials;
  const DTexture* textures;
  const DMedium* media;
};
RTB_DEV Tables scene_tables(const DScene& S) {
  Tables T;
  T.materials = S.materials; T.textures = S.textures; T.media = S.media;
  return T;
}

// ------------------------------------------------------------------------------------------------
// primitive tests, f64.  Upper bound CLOSED in both (the caller applies the reference's
// tie rule); lower bound closed for quads (Interval::contains) and open for spheres
// (Interval::surrounds) -- src/object.rs:161-163, 462, src/interval.rs:21-27 (Q4).
// ------------------------------------------------------------------------------------------------
// Quad::hit  src/object.rs:453-490, operation by operation (payload: normal d | q | u | v | w)
RTB_DEV bool quad_test(const double2* __restrict__ P, const double2 n01, const double2 n2d, const Ray& r, double tmin,
                       double tmax, double& t_out, double& a_out, double& b_out) {
  // n01, n2d = P[0], P[1], loaded by the caller (test_prim issues them together with prim_info)
  // (256-bit loads of the payload -- one instruction per 32-byte sector, as the node loads of the wavefront
  // traversal use -- were measured here too: 1 % slower, the wider destination costs two registers.)
  const double denom = ddot(n01.x, n01.y, n2d.x, r.dx, r.dy, r.dz);
  if (fabs(denom) < 1e-8) return false;
  const double t = dsub(n2d.y, ddot(n01.x, n01.y, n2d.x, r.ox, r.oy, r.oz)) / denom;
  if (!(tmin <= t && t <= tmax)) return false;
  const double2 q01 = RTB_LDG(P + 2), q2u0 = RTB_LDG(P + 3), u12 = RTB_LDG(P + 4);
  const double2 v01 = RTB_LDG(P + 5), v2w0 = RTB_LDG(P + 6), w12 = RTB_LDG(P + 7);
  // intersection = r.at(t); planar_hitpt_vector = intersection - q
  const double hx = dsub(dadd(r.ox, dmul(t, r.dx)), q01.x);
  const double hy = dsub(dadd(r.oy, dmul(t, r.dy)), q01.y);
  const double hz = dsub(dadd(r.oz, dmul(t, r.dz)), q2u0.x);
  const double ux = q2u0.y, uy = u12.x, uz = u12.y, vx = v01.x, vy = v01.y, vz = v2w0.x;
  const double wx = v2w0.y, wy = w12.x, wz = w12.y;
  // a = dot(w, cross(hp, v))
  const double a = ddot(wx, wy, wz, dsub(dmul(hy, vz), dmul(hz, vy)), dsub(dmul(hz, vx), dmul(hx, vz)),
                        dsub(dmul(hx, vy), dmul(hy, vx)));
  if ((a < 0.) || (1. < a)) return false;
  // b = dot(w, cross(u, hp))
  const double b = ddot(wx, wy, wz, dsub(dmul(uy, hz), dmul(uz, hy)), dsub(dmul(uz, hx), dmul(ux, hz)),
                        dsub(dmul(ux, hy), dmul(uy, hx)));
  if ((b < 0.) || (1. < b)) return false;
  t_out = t; a_out = a; b_out = b;
  return true;
}
RTB_DEV bool quad_test(const double2* __restrict__ P, const Ray& r, double tmin, double tmax, double& t_out,
                       double& a_out, double& b_out) {
  return quad_test(P, RTB_LDG(P + 0), RTB_LDG(P + 1), r, tmin, tmax, t_out, a_out, b_out);
}

// Sphere::hit  src/object.rs:145-166, operation by operation (root selection only; normal/uv are
// completed for the winner).  payload: cx cy | cz r | cvx cvy | cvz -
RTB_DEV bool sphere_test(const double2* __restrict__ P, const double2 c01, const double2 c2r, int moving, const Ray& r,
                         double time, double tmin, double tmax, double& t_out) {
  double cx = c01.x, cy = c01.y, cz = c2r.x;
  if (moving) {  // Sphere::center  src/object.rs:107-112: self.center + time * dir
    const double2 v01 = RTB_LDG(P + 2), v2 = RTB_LDG(P + 3);
    cx = dadd(cx, dmul(time, v01.x)); cy = dadd(cy, dmul(time, v01.y)); cz = dadd(cz, dmul(time, v2.x));
  }
  const double ocx = dsub(r.ox, cx), ocy = dsub(r.oy, cy), ocz = dsub(r.oz, cz);
  const double a = ddot(r.dx, r.dy, r.dz, r.dx, r.dy, r.dz);
  const double half_b = ddot(ocx, ocy, ocz, r.dx, r.dy, r.dz);
  const double c = dsub(ddot(ocx, ocy, ocz, ocx, ocy, ocz), dmul(c2r.y, c2r.y));
  const double disc = dsub(dmul(half_b, half_b), dmul(a, c));
  if (disc < 0.) return false;
  const double sqrtd = sqrt(disc);
  double root = dsub(-half_b, sqrtd) / a;
  if (!(tmin < root && root <= tmax)) {
    if (root > tmax) return false;  // the far root is even larger
    root = dsub(sqrtd, half_b) / a;
    if (!(tmin < root && root <= tmax)) return false;
  }
  t_out = root;
  return true;
}
RTB_DEV bool sphere_test(const double2* __restrict__ P, int moving, const Ray& r, double time, double tmin, double tmax,
                         double& t_out) {
  return sphere_test(P, RTB_LDG(P + 0), RTB_LDG(P + 1), moving, r, time, tmin, tmax, t_out);
}

// Tie rule equivalent to HittableList::hit's in-order scan (src/hittable.rs:92-106, Q7): a later
// quad replaces an equal-t hit (closed interval), a later sphere does not (open interval).
RTB_DEV bool tie_wins(int kind, int id, int best_kind, int best_id) {
  if (kind == PRIM_QUAD) return best_kind != PRIM_QUAD || id > best_id;
  return best_kind != PRIM_QUAD && id < best_id;
}

struct Hit {
  double t;     // +inf: none
  double a, b;  // quad planar coordinates of the winner
  int prim;     // index into prims (BVH order); -1 none
};

RTB_DEV void hit_reset(Hit& h) { h.t = RTB_INF; h.a = 0.; h.b = 0.; h.prim = -1; }

// a primitive test passed with t <= best.t: strictly closer wins; an exact tie is decided by the
// reference's in-order scan (tie_wins) from the two primitives' kinds and canonical ids
RTB_DEV void accept_hit(const DScene& S, int pi, int kind, double t, double a, double b, Hit& best) {
  bool take = t < best.t;
  if (!take) {
    const int4 me = RTB_LDG(S.prim_info + pi), other = RTB_LDG(S.prim_info + best.prim);
    take = tie_wins(kind, me.w, other.x & 0xFF, other.w);
  }
  if (take) { best.t = t; best.a = a; best.b = b; best.prim = pi; }
}

RTB_DEV void test_prim(const DScene& S, int pi, const Ray& r, double tmin, Hit& best) {
  RTB_ASSERT(pi >= 0 && pi < S.n_prims);
  const double2* P = S.prims + (size_t)pi * PRIM_D2;
  // the first 32 payload bytes are needed whatever the kind: issue them together with prim_info
  // instead of behind the kind branch (one load latency less on the leaf path)
  const int4 info = RTB_LDG(S.prim_info + pi);
  const double2 p0 = RTB_LDG(P + 0), p1 = RTB_LDG(P + 1);
  const int kind = info.x & 0xFF;
  double t, a = 0., b = 0.;
  bool hit;
  if (kind == PRIM_QUAD) hit = quad_test(P, p0, p1, r, tmin, best.t, t, a, b);
  else hit = sphere_test(P, p0, p1, info.x & PRIM_FLAG_MOVING, r, r.time, tmin, best.t, t);
  if (hit) accept_hit(S, pi, kind, t, a, b, best);
}

// the same test for the single primitive of a leaf whose reference carries the kind bits: no prim_info read
RTB_DEV void test_prim_k(const DScene& S, int pi, int kind_bits, const Ray& r, double tmin, Hit& best) {
  RTB_ASSERT(pi >= 0 && pi < S.n_prims);
  const double2* P = S.prims + (size_t)pi * PRIM_D2;
  const double2 p0 = RTB_LDG(P + 0), p1 = RTB_LDG(P + 1);
  double t, a = 0., b = 0.;
  bool hit;
  if (kind_bits & LEAF_KIND_QUAD) hit = quad_test(P, p0, p1, r, tmin, best.t, t, a, b);
  else hit = sphere_test(P, p0, p1, kind_bits & LEAF_KIND_MOVING, r, r.time, tmin, best.t, t);
  if (hit) accept_hit(S, pi, (kind_bits & LEAF_KIND_QUAD) ? PRIM_QUAD : PRIM_SPHERE, t, a, b, best);
}

// all primitives of a leaf.  MULTI = false: the builder made one-primitive leaves only (its default), the
// generic loop -- a second inlined copy of both primitive tests -- is compiled out (extend: -1.1 %).
template <bool MULTI = true>
RTB_DEV int test_leaf(const DScene& S, int leaf_ref, const Ray& r, double tmin, Hit& best) {
  const int first = leaf_first(leaf_ref), count = leaf_count(leaf_ref);
  if (leaf_kind_bits(leaf_ref) == LEAF_KIND_BOX) {  // an axis-aligned make_box: its six quads, as the reference's List scans them
    for (int i = 0; i < 6; i++) test_prim_k(S, first + i, LEAF_KIND_QUAD, r, tmin, best);
  } else if (!MULTI || count == 1) {
    test_prim_k(S, first, leaf_kind_bits(leaf_ref), r, tmin, best);
  } else {
    for (int i = 0; i < count; i++) test_prim(S, first + i, r, tmin, best);
  }
  return count;
}

// ------------------------------------------------------------------------------------------------
// One BVH2 node = both children's boxes (4 x float4):
//   n0 = lo0x hi0x lo0y hi0y | n1 = lo1x hi1x lo1y hi1y | n2 = lo0z hi0z lo1z hi1z | n3 = child refs
// Slab distances in fused form: t = plane * inv_d - o * inv_d (one FFMA per plane).
// Conservative: boxes are padded on the host, far gets a 4e-6 relative slack, NaNs are dropped by
// fminf/fmaxf, which only ever widens the interval.
// A/B arm RTB_SLAB_CENTER (measured, no gain: 62.9 vs 62.4 ms per c4 step): centre / half-extent
// slots, t_mid = c * inv_d - o * inv_d, near/far = t_mid -/+ h |inv_d| -- trades the 12 per-axis
// FMNMX of a node for 6 FFMA, but the loop is not bound by the ALU pipe alone.
// ------------------------------------------------------------------------------------------------
struct SlabRay { float idx, idy, idz, oxi, oyi, ozi; };

// Exactly-zero direction components DO occur (r2 = 0 in the cosine map returns the surface normal,
// e.g. (0,1,0) off the ground; about a dozen rays per 64 M paths): with 1/0 = inf the centre/extent
// form turns those axes into NaN = "unconstrained" and the ray visits most of the tree (measured:
// single rays stretching an extend launch from 0.47 to 3.4 ms).  A tiny non-zero stand-in keeps the
// reciprocal finite; the slab along that axis then culls correctly and conservatively.
RTB_DEV float fast_rcp(float x) {  // MUFU.RCP (<= 1 ulp, 1/0 = inf as IEEE) -- only ever used by conservative culls
#if defined(__CUDACC__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
// fp32 shading arithmetic (sampling maps, optics, pdf weights) is held to the oracle statistically, not bit by
// bit: MUFU.SQRT / MUFU.RCP (<= 2 ulp) instead of the IEEE sequences with their slow-path branches.
RTB_DEV float fast_sqrt(float x) {
#if defined(__CUDACC__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
RTB_DEV float fast_div(float a, float b) {
#if defined(__CUDACC__)
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
// The reciprocal only feeds the conservative cull (4e-6 relative slack): MUFU.RCP (<= 1 ulp) instead of the IEEE
// division sequence, whose slow path showed up with 10 % of the extend kernel's stall samples.
RTB_DEV float safe_rcp(float d) {
  return fast_rcp(fabsf(d) >= 1e-20f ? d : copysignf(1e-20f, d));
}

RTB_DEV SlabRay slab_ray(double ox, double oy, double oz, float dx, float dy, float dz) {
  SlabRay s;
  s.idx = safe_rcp(dx); s.idy = safe_rcp(dy); s.idz = safe_rcp(dz);
  s.oxi = (float)ox * s.idx; s.oyi = (float)oy * s.idy; s.ozi = (float)oz * s.idz;
  return s;
}

RTB_DEV void slab_box(float cx, float hx, float cy, float hy, float cz, float hz, const SlabRay& s, float tmin32,
                      float tbest32, float& tn, bool& hit) {
#if !defined(RTB_SLAB_CENTER)
  const float a0 = fmaf(cx, s.idx, -s.oxi), a1 = fmaf(hx, s.idx, -s.oxi);
  const float b0 = fmaf(cy, s.idy, -s.oyi), b1 = fmaf(hy, s.idy, -s.oyi);
  const float c0 = fmaf(cz, s.idz, -s.ozi), c1 = fmaf(hz, s.idz, -s.ozi);
  tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), tmin32));
  const float tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), tbest32));
#else
  // A/B arm: the slots hold centre / half-extent instead of lo / hi
  const float mx = fmaf(cx, s.idx, -s.oxi), my = fmaf(cy, s.idy, -s.oyi), mz = fmaf(cz, s.idz, -s.ozi);
  const float ex = hx * fabsf(s.idx), ey = hy * fabsf(s.idy), ez = hz * fabsf(s.idz);
  tn = fmaxf(fmaxf(mx - ex, my - ey), fmaxf(mz - ez, tmin32));
  const float tf = fminf(fminf(mx + ex, my + ey), fminf(mz + ez, tbest32));
#endif
  hit = tn <= fmaf(fabsf(tf), 4e-6f, tf);
}

// ------------------------------------------------------------------------------------------------
// 32-byte nodes (DScene::qnodes): the same BVH2 with child boxes as 16-bit cell indices of a scene-
// wide grid (lo | hi << 16 per axis; 2 x uint4 = {x, y, z, child ref} per child).  Why: ncu on the
// traversal shows the L1 data pipe at 75 % of its peak -- every load instruction of a warp costs one
// wavefront per distinct 32 B sector, lanes sit on different nodes, and a 64-byte fp32 node takes 4
// load instructions.  A 32-byte node takes 2 and fills exactly one sector.
// Decode without integer conversion: PRMT places the 16-bit index k in the mantissa of 2^23
// (0x4B00kkkk = 2^23 + k exactly), and t = (2^23 + k) * inv_d - (2^23 + o') * inv_d is ONE FFMA; the
// rounded second product moves a plane by <= 1 cell, which the builder's outward padding of 2 cells
// absorbs (flatten.cpp).  Rays whose origin is more than 2^23 cells from the grid traverse unculled.
// ------------------------------------------------------------------------------------------------
RTB_DEV SlabRay slab_ray_q(const DScene& S, double ox, double oy, double oz, float dx, float dy, float dz) {
  SlabRay s;
  const double gx = (ox - S.grid_base[0]) * S.grid_inv_cell[0];
  const double gy = (oy - S.grid_base[1]) * S.grid_inv_cell[1];
  const double gz = (oz - S.grid_base[2]) * S.grid_inv_cell[2];
  s.idx = S.grid_cell[0] * safe_rcp(dx); s.idy = S.grid_cell[1] * safe_rcp(dy); s.idz = S.grid_cell[2] * safe_rcp(dz);
  s.oxi = (float)((8388608.0 + gx) * (double)s.idx);
  s.oyi = (float)((8388608.0 + gy) * (double)s.idy);
  s.ozi = (float)((8388608.0 + gz) * (double)s.idz);
  if (!(fabs(gx) <= 8388608.0 && fabs(gy) <= 8388608.0 && fabs(gz) <= 8388608.0)) {
    // far outside the grid (or non-finite): NaN slab distances are dropped by fminf/fmaxf -> no culling
    s.idx = s.idy = s.idz = __int_as_float(0x7FC00000);
  }
  return s;
}

RTB_DEV void slab_box_q(uint32_t qx, uint32_t qy, uint32_t qz, const SlabRay& s, float tmin32, float tbest32, float& tn,
                        bool& hit) {
  const uint32_t M = 0x4B000000u;  // 2^23
  const float a0 = fmaf(__uint_as_float(__byte_perm(qx, M, 0x7610)), s.idx, -s.oxi);
  const float a1 = fmaf(__uint_as_float(__byte_perm(qx, M, 0x7632)), s.idx, -s.oxi);
  const float b0 = fmaf(__uint_as_float(__byte_perm(qy, M, 0x7610)), s.idy, -s.oyi);
  const float b1 = fmaf(__uint_as_float(__byte_perm(qy, M, 0x7632)), s.idy, -s.oyi);
  const float c0 = fmaf(__uint_as_float(__byte_perm(qz, M, 0x7610)), s.idz, -s.ozi);
  const float c1 = fmaf(__uint_as_float(__byte_perm(qz, M, 0x7632)), s.idz, -s.ozi);
  tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), tmin32));
  const float tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), tbest32));
  hit = tn <= fmaf(fabsf(tf), 4e-6f, tf);
}

// ------------------------------------------------------------------------------------------------
// closest surface hit: fp32 BVH2 cull + f64 leaves.  Replaces HittableList::hit / BvhNode::hit /
// Aabb::hit (src/hittable.rs:88-109, 216-236, src/object.rs:340-370) with the build's own tree.
// ------------------------------------------------------------------------------------------------
template <bool STATS>
RTB_DEV void closest_surface(const DScene& S, const Ray& r, double tmin, Hit& best, DStats* st) {
  const SlabRay sr = slab_ray(r.ox, r.oy, r.oz, (float)r.dx, (float)r.dy, (float)r.dz);
  const float tmin32 = __double2float_rd(tmin);
  float tbest32 = __double2float_ru(best.t);
  int stack[BVH_STACK];
  int sp = 0;
  int node = 0;
  RTB_LOOP_ENTER();
  for (;;) {
    if (node >= 0) {
      if (STATS) st->node_visits++;
      RTB_ASSERT(node < S.n_nodes);
      const float4* N = S.nodes + 4 * (size_t)node;
      const float4 n0 = RTB_LDG(N + 0), n1 = RTB_LDG(N + 1), n2 = RTB_LDG(N + 2), n3 = RTB_LDG(N + 3);
      float tn0, tn1;
      bool h0, h1;
      slab_box(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, sr, tmin32, tbest32, tn0, h0);
      slab_box(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, sr, tmin32, tbest32, tn1, h1);
      int ch0 = __float_as_int(n3.x), ch1 = __float_as_int(n3.y);
      if (h0 && h1) {
        if (tn1 < tn0) { const int tmp = ch0; ch0 = ch1; ch1 = tmp; }
        RTB_ASSERT(sp < BVH_STACK);
        stack[sp++] = ch1;
        node = ch0;
        continue;
      }
      if (h0) { node = ch0; continue; }
      if (h1) { node = ch1; continue; }
    } else {
      const int count = test_leaf(S, node, r, tmin, best);
      if (STATS) st->prim_tests += (unsigned long long)count;
      tbest32 = __double2float_ru(best.t);
    }
    if (sp == 0) break;
    node = stack[--sp];
  }
  RTB_LOOP_LEAVE();
}

// linear scan over all surface primitives (RTB_TRACE_BRUTE_FORCE: validates the BVH cull)
RTB_DEV void closest_surface_brute(const DScene& S, const Ray& r, double tmin, Hit& best) {
  for (int i = 0; i < S.n_surface_prims; i++) test_prim(S, i, r, tmin, best);
}

// ------------------------------------------------------------------------------------------------
// Candidate scheme of the wavefront pipeline: the traversal decides NOTHING that could change a result.
//
// The f64 reference-order tests above are what holds first-hit ids bit-exact, but inside a traversing warp
// they run at 5-8 of 32 lanes behind a division and a square root (ncu, profiles/r01b_k_wf_extend_*).  So
// the traversal only classifies each leaf primitive with a cheap test whose every rounding error is bounded:
//     PF_MISS    the reference test certainly rejects it (or it certainly lies behind a certain hit),
//     PF_HIT     the reference test certainly accepts it, with its t inside [t_lo, t_hi],
//     PF_UNSURE  anything else (edges, grazing rays, near-parallel planes, t_min straddled),
// keeps the (at most CAND_K) primitives that can still be the closest hit, and the shade stage runs the exact
// test on those survivors with every lane busy.  More live candidates than slots -> the ray is re-traced by
// the exact kernel (k_wf_extend_exact).  Same closest hit as closest_surface(), bit for bit: a primitive is
// only ever dropped when its hit is certainly absent or certainly farther than another certain hit.
//
// Arithmetic: fp32 with running error bounds (u = 2^-24; the constants carry >= 2x headroom over the
// first-order terms, and every "certain" decision is a positive comparison, so NaN/inf fall to UNSURE);
// only the differences that cancel -- plane distance d - n.o, o - q, o - c, |o-c|^2 - r^2 -- are f64
// (7 DFMA-class operations per test instead of ~50 plus a division and a square root).  Directions may
// carry a rounding of their own (primary rays are f64 in the queue): 1 u per component, inside the bounds.
// ------------------------------------------------------------------------------------------------
enum : int { PF_MISS = 0, PF_HIT = 1, PF_UNSURE = 2 };
constexpr float PF_U = 5.9604645e-8f;  // 2^-24
constexpr int CAND_K = 2;
constexpr int CAND_OVERFLOW = 1;  // candidate-slot sentinel: "re-trace me exactly" (leaf references are negative)
constexpr int CAND_CERTAIN = 2;   // in the SECOND slot: the first is the only candidate and a certain hit -- only its t is open

struct PfRay {
  double ox, oy, oz;
  float dx, dy, dz, time;
  float d1;       // |dx| + |dy| + |dz|
  float dd;       // d . d
  float num_err;  // bound of the f64 rounding of d - n.o and o - q for this origin (4e-13 x coordinate magnitude)
};
// scene_mag bounds every coordinate in play (primitives, camera, and -- for the parity harness, whose rays may start
// anywhere -- the ray origins of the call): |d| + sum |n_i o_i| <= 4 scene_mag, times 500 x the f64 unit roundoff
RTB_DEV PfRay pf_ray(double ox, double oy, double oz, float dx, float dy, float dz, float time, float scene_mag) {
  PfRay r;
  r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz; r.time = time;
  r.d1 = fabsf(dx) + fabsf(dy) + fabsf(dz);
  r.dd = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
  r.num_err = 2e-12f * scene_mag;
  return r;
}

// 32 bytes of a primitive payload / prefilter record in one load instruction (one L1 wavefront per sector)
struct D4 { double a, b, c, d; };
RTB_DEV D4 load_d4(const double2* p) {
#if defined(__CUDACC__)
  D4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
  return r;
#else
  D4 r;
  r.a = p[0].x; r.b = p[0].y; r.c = p[1].x; r.d = p[1].y;
  return r;
#endif
}
// 256-bit global load (sm_100: LDG.E.ENL2.256): one instruction per 32-byte sector.  The L1 data pipe
// charges a wavefront per load instruction and distinct sector, and the lanes of a traversing warp sit on
// different nodes -- so a 64-byte node read as 2 x 256 bit costs half the wavefronts of 4 x 128 bit.
struct alignas(32) F8 { float4 lo, hi; };
RTB_DEV F8 load_f8(const float4* p) {
  F8 r;
#if defined(__CUDACC__)
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p));
#else
  r.lo = p[0]; r.hi = p[1];
#endif
  return r;
}

// Quad::hit (src/object.rs:453-490), conservatively.  alpha = w.(h x v) = h.(v x w) =: h.A and
// beta = w.(u x h) = h.(w x u) =: h.B with A, B from the flattener (DPre); h = (o - q) + t d.
RTB_DEV int prefilter_quad(const double2* __restrict__ P, const DPre* __restrict__ pre, const PfRay& r, float tmin_lo,
                           float tmin_hi, float bound, float& t_lo, float& t_hi) {
  const D4 nd = load_d4(P);  // unit normal, d = n.q
#if defined(__CUDACC__)
  const D4 q4 = load_d4(reinterpret_cast<const double2*>(pre));
  const F8 f = load_f8(reinterpret_cast<const float4*>(pre) + 2);
  const double qx = q4.a, qy = q4.b, qz = q4.c;
  const float ax = __int_as_float(__double2loint(q4.d)), ay = __int_as_float(__double2hiint(q4.d));
  const float az = f.lo.x, bx = f.lo.y, by = f.lo.z, bz = f.lo.w, nx = f.hi.x, ny = f.hi.y, nz = f.hi.z, ab1 = f.hi.w;
#else
  const double qx = pre->qx, qy = pre->qy, qz = pre->qz;
  const float ax = pre->ax, ay = pre->ay, az = pre->az, bx = pre->bx, by = pre->by, bz = pre->bz;
  const float nx = pre->nx, ny = pre->ny, nz = pre->nz, ab1 = pre->ab1;
#endif
  t_lo = tmin_lo; t_hi = 0.f;
  const float den = fmaf(nx, r.dx, fmaf(ny, r.dy, nz * r.dz));
  const float den_err = 10.f * PF_U * fmaf(fabsf(nx), fabsf(r.dx), fmaf(fabsf(ny), fabsf(r.dy), fabsf(nz * r.dz)));
  // near-parallel: the reference's |denom| < 1e-8 rule decides, exactly
  if (!(fabsf(den) > 1e-7f + 4.f * den_err)) return PF_UNSURE;
  const float num = (float)(nd.d - fma(nd.c, r.oz, fma(nd.b, r.oy, nd.a * r.ox)));
  const float inv = fast_rcp(den);
  const float t = num * inv;
  const float tw = fmaf(fabsf(t), fmaf(1.5f * den_err, fabsf(inv), 10.f * PF_U), 1.5f * r.num_err * fabsf(inv)) + 1e-30f;
  t_lo = t - tw; t_hi = t + tw;
  if (t_hi < tmin_lo || t_lo > bound) return PF_MISS;
  const float ex = (float)(r.ox - qx), ey = (float)(r.oy - qy), ez = (float)(r.oz - qz);
  const float hx = fmaf(t, r.dx, ex), hy = fmaf(t, r.dy, ey), hz = fmaf(t, r.dz, ez);
  const float h_err = fmaf(4.f * PF_U, fabsf(ex) + fabsf(ey) + fabsf(ez) + fabsf(t) * r.d1, tw * r.d1) + r.num_err;
  const float ab_err = ab1 * fmaf(8.f * PF_U, fabsf(hx) + fabsf(hy) + fabsf(hz), h_err) + 4.f * PF_U;
  const float a = fmaf(ax, hx, fmaf(ay, hy, az * hz));
  const float b = fmaf(bx, hx, fmaf(by, hy, bz * hz));
  if (a < -ab_err || a > 1.f + ab_err || b < -ab_err || b > 1.f + ab_err) return PF_MISS;
  const bool inside = a > ab_err && a < 1.f - ab_err && b > ab_err && b < 1.f - ab_err;
  return (inside && t_lo >= tmin_hi) ? PF_HIT : PF_UNSURE;
}

// Sphere::hit (src/object.rs:145-166), conservatively.  Roots through q = -(hb + sgn(hb) sqrt(disc)), q / a
// and c / q: no cancellation, so a ray leaving its own sphere sees the root at 0 as ~1e-13, not as ~u |hb|.
// An UNSURE result carries the smallest t the reference could return in t_lo (used for pruning only).
RTB_DEV int prefilter_sphere(const double2* __restrict__ P, bool moving, const PfRay& r, float tmin_lo, float tmin_hi,
                             float bound, float& t_lo, float& t_hi) {
  const D4 cr = load_d4(P);
  double cx = cr.a, cy = cr.b, cz = cr.c;
  if (moving) {  // Sphere::center, the reference's own two roundings (src/object.rs:107-112)
    const D4 cv = load_d4(P + 2);
    const double tm = (double)r.time;
    cx = dadd(cx, dmul(tm, cv.a)); cy = dadd(cy, dmul(tm, cv.b)); cz = dadd(cz, dmul(tm, cv.c));
  }
  t_lo = tmin_lo; t_hi = 0.f;
  const double ocx64 = r.ox - cx, ocy64 = r.oy - cy, ocz64 = r.oz - cz;
  const float ocx = (float)ocx64, ocy = (float)ocy64, ocz = (float)ocz64;
  const float cc = (float)fma(-cr.d, cr.d, fma(ocz64, ocz64, fma(ocy64, ocy64, ocx64 * ocx64)));
  // f64 rounding of |oc|^2 - r^2 (the reference's and ours): 4e-15 (|oc|^2 + r^2), with r^2 = |oc|^2 - cc
  const float cc_err = 4e-15f * fmaf(2.1f, fmaf(ocx, ocx, fmaf(ocy, ocy, ocz * ocz)), fabsf(cc)) + 1e-30f;
  const float hb = fmaf(ocx, r.dx, fmaf(ocy, r.dy, ocz * r.dz));
  const float hb_err = 10.f * PF_U * fmaf(fabsf(ocx), fabsf(r.dx), fmaf(fabsf(ocy), fabsf(r.dy), fabsf(ocz * r.dz))) + 1e-30f;
  const float ac = r.dd * cc;
  const float disc = fmaf(hb, hb, -ac);
  const float disc_err = fmaf(2.f * fabsf(hb) + hb_err, hb_err, fmaf(12.f * PF_U, fmaf(hb, hb, fabsf(ac)), r.dd * cc_err));
  if (disc < -disc_err) return PF_MISS;
  if (!(disc > 4.f * disc_err)) return PF_UNSURE;  // grazing
  const float sq = fast_sqrt(disc);
  const float sq_err = fmaf(0.6f * disc_err, fast_rcp(sq), 4.f * PF_U * sq);
  const float q = -(hb + copysignf(sq, hb));
  const float q_err = hb_err + sq_err + 2.f * PF_U * fabsf(q);
  const float inv_q = fast_rcp(q);
  const float rq = q_err * fabsf(inv_q);
  if (!(rq < 0.25f)) return PF_UNSURE;
  const float ra = q * fast_rcp(r.dd), rb = cc * inv_q;
  const float wa = fabsf(ra) * fmaf(1.4f, rq, 12.f * PF_U) + 1e-30f;
  const float wb = fmaf(fabsf(cc), fmaf(1.4f, rq, 12.f * PF_U), 1.4f * cc_err) * fabsf(inv_q) + 1e-30f;
  const bool neg = q > 0.f;  // hb < 0 (or -0): ra is the FAR root
  const float r1 = neg ? rb : ra, w1 = neg ? wb : wa, r2 = neg ? ra : rb, w2 = neg ? wa : wb;
  // reference root selection over (t_min, closest]: the near root if it lies beyond t_min, else the far one
  if (r1 - w1 > tmin_hi) {
    t_lo = r1 - w1; t_hi = r1 + w1;
    return t_lo > bound ? PF_MISS : PF_HIT;
  }
  if (r1 + w1 < tmin_lo) {  // near root certainly rejected
    if (r2 - w2 > tmin_hi) {
      t_lo = r2 - w2; t_hi = r2 + w2;
      return t_lo > bound ? PF_MISS : PF_HIT;
    }
    if (r2 + w2 < tmin_lo) return PF_MISS;
    t_lo = r2 - w2;
    return t_lo > bound ? PF_MISS : PF_UNSURE;
  }
  t_lo = r1 - w1;  // t_min straddled: the reference returns r1, r2 or nothing
  return t_lo > bound ? PF_MISS : PF_UNSURE;
}

// An axis-aligned make_box (src/object.rs:509-560: six quads) as ONE leaf.  Its faces are axis-aligned, so the
// reference's Quad::hit of face (k, side) evaluates t = (plane_k - o_k) / d_k -- the normal is (+-1, 0, 0) exactly and the
// zero terms of both dot products vanish -- and its alpha / beta test is "the hit point lies within the other two
// slabs".  One slab test therefore names the only face that can be the closest hit: the ENTRY face when the box is
// entered beyond t_min, the EXIT face when the origin lies inside (or on a face, heading in).  Every decision carries its
// error window; anything ambiguous (edges, corners, grazing, t_min straddled, near-parallel entry) gives up:
//   returns PF_MISS (no face can be the closest hit), PF_HIT (face index in `face`, t in [t_lo, t_hi]), or PF_UNSURE.
RTB_DEV int prefilter_box(const DBoxBounds* __restrict__ bb, const PfRay& r, float tmin_lo, float tmin_hi, float bound, int& face,
                          float& t_lo, float& t_hi) {
  const D4 b0 = load_d4(reinterpret_cast<const double2*>(bb)), b1 = load_d4(reinterpret_cast<const double2*>(bb) + 2);
  // (plane - o) in f64: a ray leaving a face of this very box must see that face at t ~ 1e-13, not ~ u |o| / |d|
  float ax = (float)(b0.a - r.ox), ay = (float)(b0.b - r.oy), az = (float)(b0.c - r.oz);
  float bx = (float)(b0.d - r.ox), by = (float)(b1.a - r.oy), bz = (float)(b1.b - r.oz);
  float ix = fast_rcp(r.dx), iy = fast_rcp(r.dy), iz = fast_rcp(r.dz);
  if (!(fabsf(r.dx) > 1e-7f && fabsf(r.dy) > 1e-7f && fabsf(r.dz) > 1e-7f)) {
    // A direction component too small to divide by is left to the exact tests -- unless it is exactly zero: that
    // coordinate never changes, the faces of that axis see denom == 0 (no hit), and the faces of the other axes are hit
    // within their extent only if the origin lies between the two planes.  Strictly outside: a miss.  Strictly inside:
    // the slab pair constrains nothing (planes at -+1e30 keep the arithmetic below finite).
    const float e = r.num_err;
    auto parallel = [e](float d, float& a, float& b, float& inv) -> int {
      if (fabsf(d) > 1e-7f) return PF_HIT;
      if (d != 0.f) return PF_UNSURE;
      if (a > e || b < -e) return PF_MISS;
      if (!(a < -e && b > e)) return PF_UNSURE;
      a = -1e30f; b = 1e30f; inv = 1.f;
      return PF_HIT;
    };
    const int cx = parallel(r.dx, ax, bx, ix), cy = parallel(r.dy, ay, by, iy), cz = parallel(r.dz, az, bz, iz);
    if (cx == PF_MISS || cy == PF_MISS || cz == PF_MISS) return PF_MISS;
    if (cx == PF_UNSURE || cy == PF_UNSURE || cz == PF_UNSURE) return PF_UNSURE;
  }
  // near / far plane of each slab pair by the sign of the direction, each distance with its own window: 6 u relative
  // (conversion, reciprocal, product, direction rounding) + the f64 rounding of the difference
  const bool px = r.dx >= 0.f, py = r.dy >= 0.f, pz = r.dz >= 0.f;
  const float n0 = (px ? ax : bx) * ix, f0 = (px ? bx : ax) * ix;
  const float n1 = (py ? ay : by) * iy, f1 = (py ? by : ay) * iy;
  const float n2 = (pz ? az : bz) * iz, f2 = (pz ? bz : az) * iz;
  const float e0 = r.num_err * fabsf(ix) + 1e-30f, e1 = r.num_err * fabsf(iy) + 1e-30f, e2 = r.num_err * fabsf(iz) + 1e-30f;
  const float wn0 = fmaf(6.f * PF_U, fabsf(n0), e0), wf0 = fmaf(6.f * PF_U, fabsf(f0), e0);
  const float wn1 = fmaf(6.f * PF_U, fabsf(n1), e1), wf1 = fmaf(6.f * PF_U, fabsf(f1), e1);
  const float wn2 = fmaf(6.f * PF_U, fabsf(n2), e2), wf2 = fmaf(6.f * PF_U, fabsf(f2), e2);
  const float nlo0 = n0 - wn0, nlo1 = n1 - wn1, nlo2 = n2 - wn2, fhi0 = f0 + wf0, fhi1 = f1 + wf1, fhi2 = f2 + wf2;
  const float tn_lo = fmaxf(nlo0, fmaxf(nlo1, nlo2)), tn_hi = fmaxf(n0 + wn0, fmaxf(n1 + wn1, n2 + wn2));
  const float tf_hi = fminf(fhi0, fminf(fhi1, fhi2));
  if (tn_lo > tf_hi) return PF_MISS;               // the line misses the box
  if (tf_hi < tmin_lo) return PF_MISS;             // the box lies behind t_min
  if (tn_lo > bound) return PF_MISS;               // ... or beyond a certain hit
  // strictly inside the other two slabs at the hit: margins of one more window (a position margin >= num_err)
  const float nin0 = fmaf(2.f, wn0, n0), nin1 = fmaf(2.f, wn1, n1), nin2 = fmaf(2.f, wn2, n2);
  const float fin0 = fmaf(-2.f, wf0, f0), fin1 = fmaf(-2.f, wf1, f1), fin2 = fmaf(-2.f, wf2, f2);
  if (tn_lo > tmin_hi) {                           // entered beyond t_min: the entry face, if it is unmistakable
    const int k = (nlo0 >= nlo1 && nlo0 >= nlo2) ? 0 : (nlo1 >= nlo2 ? 1 : 2);
    const float lo = k == 0 ? nlo0 : (k == 1 ? nlo1 : nlo2), hi = k == 0 ? n0 + wn0 : (k == 1 ? n1 + wn1 : n2 + wn2);
    const float other_near = k == 0 ? fmaxf(nin1, nin2) : (k == 1 ? fmaxf(nin0, nin2) : fmaxf(nin0, nin1));
    const float other_far = k == 0 ? fminf(fin1, fin2) : (k == 1 ? fminf(fin0, fin2) : fminf(fin0, fin1));
    if (!(other_near < lo && hi < other_far)) return PF_UNSURE;  // an edge, a corner, or grazing
    face = 2 * k + ((k == 0 ? px : (k == 1 ? py : pz)) ? 0 : 1);  // moving towards +k enters through the low plane
    t_lo = lo; t_hi = hi;
    return PF_HIT;
  }
  if (tn_hi < tmin_lo) {                           // origin inside, or on a face heading in: the exit face
    const int k = (fhi0 <= fhi1 && fhi0 <= fhi2) ? 0 : (fhi1 <= fhi2 ? 1 : 2);
    const float hi = k == 0 ? fhi0 : (k == 1 ? fhi1 : fhi2), lo = k == 0 ? f0 - wf0 : (k == 1 ? f1 - wf1 : f2 - wf2);
    if (!(lo > tmin_hi)) return PF_UNSURE;
    const float other_near = k == 0 ? fmaxf(nin1, nin2) : (k == 1 ? fmaxf(nin0, nin2) : fmaxf(nin0, nin1));
    const float other_far = k == 0 ? fminf(fin1, fin2) : (k == 1 ? fminf(fin0, fin2) : fminf(fin0, fin1));
    if (!(other_near < lo && hi < other_far)) return PF_UNSURE;
    if (lo > bound) return PF_MISS;
    face = 2 * k + ((k == 0 ? px : (k == 1 ? py : pz)) ? 1 : 0);  // moving towards +k leaves through the high plane
    t_lo = lo; t_hi = hi;
    return PF_HIT;
  }
  return PF_UNSURE;                                // t_min straddled by the entry
}

// the (at most CAND_K) leaf references that can still hold the closest hit, each with the smallest t it could have
struct Cands {
  int c0, c1;       // 0 = empty slot; CAND_OVERFLOW in c0 = more live candidates than slots
  float lo0, lo1;   // smallest t each could have; the sign bit of the REFERENCE's complement is not used -- certainty
  float bound;      // smallest t_hi of a certain hit so far (+inf: none): nothing beyond it can be the closest hit
  int certain;      // bit 0 / 1: slot 0 / 1 holds a certain hit (PF_HIT)
};
RTB_DEV void cands_reset(Cands& C) { C.c0 = 0; C.c1 = 0; C.lo0 = 0.f; C.lo1 = 0.f; C.bound = __int_as_float(0x7F800000); C.certain = 0; }
RTB_DEV void cands_add(Cands& C, int ref, int cls, float t_lo, float t_hi) {
  if (cls == PF_MISS || C.c0 == CAND_OVERFLOW) return;
  if (cls == PF_HIT && t_hi < C.bound) {
    C.bound = t_hi;
    if (C.c1 != 0 && C.lo1 > C.bound) { C.c1 = 0; C.certain &= 1; }
    if (C.c0 != 0 && C.lo0 > C.bound) { C.c0 = C.c1; C.lo0 = C.lo1; C.c1 = 0; C.certain >>= 1; }
  }
  const int cert = cls == PF_HIT ? 1 : 0;
  if (C.c0 == 0) { C.c0 = ref; C.lo0 = t_lo; C.certain = cert; }
  else if (C.c1 == 0) { C.c1 = ref; C.lo1 = t_lo; C.certain |= cert << 1; }
  else C.c0 = CAND_OVERFLOW;
}
// the candidate record that travels to the shade stage: {c0, c1}, with CAND_CERTAIN in the empty second slot when the
// only candidate is a one-primitive leaf and a certain hit (then only its t has to be evaluated exactly)
RTB_DEV void cands_record(const Cands& C, int& r0, int& r1) {
  r0 = C.c0; r1 = C.c1;
  if (C.c0 < 0 && C.c1 == 0 && (C.certain & 1) && leaf_count(C.c0) == 1) r1 = CAND_CERTAIN;
}

// Classification of one leaf against a ray: PF_MISS / PF_HIT / PF_UNSURE with the window [t_lo, t_hi], or PF_OVERFLOW
// (a box leaf the slab test cannot decide: six possible faces do not fit two slots, the exact re-trace decides).
// ref_out = the reference that becomes the candidate: the leaf itself, or the named face of a box leaf (a one-quad
// leaf reference).  A leaf of several primitives is a candidate as a whole: smallest t_lo of its live primitives,
// certain hits still tighten.  MULTI as in test_leaf.
enum : int { PF_OVERFLOW = 3 };
template <bool MULTI = true>
RTB_DEV int prefilter_classify(const DScene& S, int leaf_ref, const PfRay& r, float tmin_lo, float tmin_hi, float bound, int& ref_out,
                               float& t_lo, float& t_hi) {
  const int first = leaf_first(leaf_ref), count = leaf_count(leaf_ref);
  ref_out = leaf_ref;
  if (leaf_kind_bits(leaf_ref) == LEAF_KIND_BOX) {
    int face = 0;
    const int cls = prefilter_box(reinterpret_cast<const DBoxBounds*>(S.pre + first), r, tmin_lo, tmin_hi, bound, face, t_lo, t_hi);
    ref_out = leaf_make(first + face, 1, LEAF_KIND_QUAD);
    return cls == PF_UNSURE ? PF_OVERFLOW : cls;
  }
  if (!MULTI || count == 1) {
    const int bits = leaf_kind_bits(leaf_ref);
    const double2* P = S.prims + (size_t)first * PRIM_D2;
    return (bits & LEAF_KIND_QUAD) ? prefilter_quad(P, S.pre + first, r, tmin_lo, tmin_hi, bound, t_lo, t_hi)
                                   : prefilter_sphere(P, (bits & LEAF_KIND_MOVING) != 0, r, tmin_lo, tmin_hi, bound, t_lo, t_hi);
  }
  int leaf_cls = PF_MISS;
  float leaf_lo = __int_as_float(0x7F800000), leaf_hi = __int_as_float(0x7F800000);
  for (int i = 0; i < count; i++) {
    const int info_x = RTB_LDG(S.prim_info + first + i).x;
    const double2* P = S.prims + (size_t)(first + i) * PRIM_D2;
    const int cls = ((info_x & 0xFF) == PRIM_QUAD) ? prefilter_quad(P, S.pre + first + i, r, tmin_lo, tmin_hi, bound, t_lo, t_hi)
                                                   : prefilter_sphere(P, (info_x & PRIM_FLAG_MOVING) != 0, r, tmin_lo, tmin_hi, bound, t_lo, t_hi);
    if (cls == PF_MISS) continue;
    leaf_lo = fminf(leaf_lo, t_lo);
    if (cls == PF_HIT) { leaf_hi = fminf(leaf_hi, t_hi); leaf_cls = PF_HIT; }
    else if (leaf_cls == PF_MISS) leaf_cls = PF_UNSURE;
  }
  t_lo = leaf_lo; t_hi = leaf_hi;
  return leaf_cls;
}
RTB_DEV void cands_apply(Cands& C, int cls, int ref, float t_lo, float t_hi) {
  if (cls == PF_OVERFLOW) C.c0 = CAND_OVERFLOW;
  else if (!(t_lo > C.bound)) cands_add(C, ref, cls, t_lo, t_hi);  // (it may have been classified against an older bound)
}

// a box leaf: the face the slab test names becomes the candidate (a one-quad leaf reference)
RTB_DEV void prefilter_box_leaf(const DScene& S, int leaf_ref, const PfRay& r, float tmin_lo, float tmin_hi, Cands& C) {
  const int first = leaf_first(leaf_ref);
  float t_lo, t_hi;
  int face = 0;
  const int cls = prefilter_box(reinterpret_cast<const DBoxBounds*>(S.pre + first), r, tmin_lo, tmin_hi, C.bound, face, t_lo, t_hi);
  if (cls == PF_HIT) cands_add(C, leaf_make(first + face, 1, LEAF_KIND_QUAD), PF_HIT, t_lo, t_hi);
  else if (cls == PF_UNSURE) C.c0 = CAND_OVERFLOW;  // (six possible faces do not fit two slots: the exact re-trace decides)
}

// classify the primitive(s) of one leaf and update the candidates (the extend kernel's leaf phase); returns the primitive
// tests it stands for (a box leaf: 1).  Written out rather than through prefilter_classify: the same statements routed
// through the classify / apply pair cost the c4 extend stage 8.6 ms per step (229.8 vs 221.2) in ptxas' hands.
template <bool MULTI = true>
RTB_DEV int prefilter_leaf(const DScene& S, int leaf_ref, const PfRay& r, float tmin_lo, float tmin_hi, Cands& C) {
  const int first = leaf_first(leaf_ref), count = leaf_count(leaf_ref);
  float t_lo, t_hi;
  if (leaf_kind_bits(leaf_ref) == LEAF_KIND_BOX) {
    prefilter_box_leaf(S, leaf_ref, r, tmin_lo, tmin_hi, C);
    return 1;
  }
  if (!MULTI || count == 1) {
    const int bits = leaf_kind_bits(leaf_ref);
    const double2* P = S.prims + (size_t)first * PRIM_D2;
    const int cls = (bits & LEAF_KIND_QUAD) ? prefilter_quad(P, S.pre + first, r, tmin_lo, tmin_hi, C.bound, t_lo, t_hi)
                                            : prefilter_sphere(P, (bits & LEAF_KIND_MOVING) != 0, r, tmin_lo, tmin_hi, C.bound, t_lo, t_hi);
    cands_add(C, leaf_ref, cls, t_lo, t_hi);
  } else {
    int ref = 0;
    const int cls = prefilter_classify<true>(S, leaf_ref, r, tmin_lo, tmin_hi, C.bound, ref, t_lo, t_hi);
    cands_add(C, leaf_ref, cls, t_lo, t_hi);
  }
  return count;
}

// exact resolution of the survivors (shade stage; every lane holds a ray): the reference-order f64 tests with the
// reference's tie rule, on the candidates only.  MULTI as in test_leaf.
// Quad::hit's t alone (src/object.rs:453-461, operation by operation): for a quad the traversal has already PROVEN to be
// hit (t inside the interval, alpha / beta inside [0,1] beyond every rounding error) the planar coordinates decide nothing
RTB_DEV double quad_t_only(const double2* __restrict__ P, const Ray& r) {
  const double2 n01 = RTB_LDG(P + 0), n2d = RTB_LDG(P + 1);
  const double denom = ddot(n01.x, n01.y, n2d.x, r.dx, r.dy, r.dz);
  return dsub(n2d.y, ddot(n01.x, n01.y, n2d.x, r.ox, r.oy, r.oz)) / denom;
}

template <bool MULTI = true>
RTB_DEV void resolve_candidates(const DScene& S, int c0, int c1, const Ray& r, double tmin, Hit& best) {
  hit_reset(best);
  if (c1 == CAND_CERTAIN && (leaf_kind_bits(c0) & LEAF_KIND_QUAD)) {
    best.prim = leaf_first(c0);
    best.t = quad_t_only(S.prims + (size_t)best.prim * PRIM_D2, r);
    return;
  }
  if (c0 < 0) test_leaf<MULTI>(S, c0, r, tmin, best);
  if (c1 < 0) test_leaf<MULTI>(S, c1, r, tmin, best);
}

// scalar form of the candidate traversal (parity harness of the host build, tools/sim): the wavefront kernel
// runs the same prefilter_leaf / cands_add inside its speculative loop.  Returns false on overflow.
template <bool STATS>
RTB_DEV bool closest_candidates(const DScene& S, const Ray& r, float scene_mag, Cands& C, DStats* st) {
  const PfRay pr = pf_ray(r.ox, r.oy, r.oz, (float)r.dx, (float)r.dy, (float)r.dz, (float)r.time, scene_mag);
  const SlabRay sr = slab_ray(r.ox, r.oy, r.oz, pr.dx, pr.dy, pr.dz);
  const float tmin_lo = __double2float_rd(0.0001), tmin_hi = __double2float_ru(0.0001);
  cands_reset(C);
  int stack[BVH_STACK];
  int sp = 0, node = 0;
  RTB_LOOP_ENTER();
  bool complete = true;
  for (;;) {
    if (node >= 0) {
      if (STATS) st->node_visits++;
      const float4* N = S.nodes + 4 * (size_t)node;
      const float4 n0 = RTB_LDG(N + 0), n1 = RTB_LDG(N + 1), n2 = RTB_LDG(N + 2), n3 = RTB_LDG(N + 3);
      float tn0, tn1;
      bool h0, h1;
      slab_box(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, sr, tmin_lo, C.bound, tn0, h0);
      slab_box(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, sr, tmin_lo, C.bound, tn1, h1);
      int ch0 = __float_as_int(n3.x), ch1 = __float_as_int(n3.y);
      if (h0 && h1) {
        if (tn1 < tn0) { const int tmp = ch0; ch0 = ch1; ch1 = tmp; }
        stack[sp++] = ch1;
        node = ch0;
        continue;
      }
      if (h0) { node = ch0; continue; }
      if (h1) { node = ch1; continue; }
    } else {
      const int count = prefilter_leaf<true>(S, node, pr, tmin_lo, tmin_hi, C);
      if (STATS) st->prim_tests += (unsigned long long)count;
      if (C.c0 == CAND_OVERFLOW) { complete = false; break; }
    }
    if (sp == 0) break;
    node = stack[--sp];
  }
  RTB_LOOP_LEAVE();
  return complete;
}

// ------------------------------------------------------------------------------------------------
// constant media.  ConstantMedium::hit  src/constant_medium.rs:41-95 (Q17), restated order-
// independently: the boundary interval comes from two probes over the medium's own boundary
// primitives; the free-flight event is accepted inside [max(t1,tmin), min(t2, t_closest)].
// ------------------------------------------------------------------------------------------------
RTB_DEV double boundary_probe(const DScene& S, const DMedium& m, const Ray& r, double tmin) {
  Hit h;
  hit_reset(h);
  for (int i = 0; i < m.n_prims; i++) test_prim(S, m.first_prim + i, r, tmin, h);
  return h.t;  // +inf: no hit
}

RTB_DEV bool medium_line_cull(const DMedium& m, const Ray& r) {  // fp32 padded box vs the whole line
  const float ox = (float)r.ox, oy = (float)r.oy, oz = (float)r.oz;
  const float idx = fast_rcp((float)r.dx), idy = fast_rcp((float)r.dy), idz = fast_rcp((float)r.dz);
  const float a0 = (m.lo[0] - ox) * idx, a1 = (m.hi[0] - ox) * idx;
  const float b0 = (m.lo[1] - oy) * idy, b1 = (m.hi[1] - oy) * idy;
  const float c0 = (m.lo[2] - oz) * idz, c1 = (m.hi[2] - oz) * idz;
  const float tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fminf(c0, c1));
  const float tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fmaxf(c0, c1));
  return tn <= fmaf(fabsf(tf), 2e-6f, tf) + 1e-30f;
}

// Boundary = an oriented box (make_box under Translate / RotateY: cls_fast bit 10).  The medium's world-space AABB is
// loose around a rotated box; in the box's own frame the slab test is tight.  fp32, conservative both ways:
//   0  the whole LINE misses the box inflated by the margin: all six Quad::hit fail, both probes return None;
//   2  both end points of [ta, tb] lie inside the box deflated by the margin (a box is convex): probe 1 over UNIVERSE
//      returns the entry behind the origin (t1 < 0 < tmin), probe 2 the exit beyond tb -- the clamps of
//      constant_medium.rs:58-63 then make the interval [tmin, tmax] whatever t1, t2 are;
//   1  anything else: evaluate the quads.
//   3  (only when `faces` is asked for) as 1, and the ENTRY and EXIT face of the line are unmistakable: the planes of the
//      true box lie between the deflated and the inflated ones, so each plane distance has an interval; the entry face is
//      named when its interval lies above the other two near intervals and below the other two far intervals, likewise
//      the exit face -- then Quad::hit succeeds on exactly these two faces (alpha / beta inside [0,1] by the margin,
//      outside it on the other four) and their two t are the two probes' results.  The boundary primitives of such a
//      medium are stored in face order (flatten.cpp): first_prim + 2 k + (the +axis_k side).
RTB_DEV int medium_obb(const DMedium& m, const Ray& r, float ta, float tb, bool faces, int& face_in, int& face_out) {
  const float ex = (float)r.ox - m.obb_c[0], ey = (float)r.oy - m.obb_c[1], ez = (float)r.oz - m.obb_c[2];
  const float dx = (float)r.dx, dy = (float)r.dy, dz = (float)r.dz;
  float tn = -3.0e38f, tf = 3.0e38f;
  bool inside = true;
  float nlo[3], nhi[3], flo[3], fhi[3], dk[3];  // (statically indexed only: registers)
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float o = fmaf(ex, m.obb_ax[k][0], fmaf(ey, m.obb_ax[k][1], ez * m.obb_ax[k][2]));
    const float d = fmaf(dx, m.obb_ax[k][0], fmaf(dy, m.obb_ax[k][1], dz * m.obb_ax[k][2]));
    const float inv = safe_rcp(d);
    const float t0 = (-m.obb_half_out[k] - o) * inv, t1 = (m.obb_half_out[k] - o) * inv;
    tn = fmaxf(tn, fminf(t0, t1));
    tf = fminf(tf, fmaxf(t0, t1));
    inside = inside && fabsf(fmaf(ta, d, o)) < m.obb_half_in[k] && fabsf(fmaf(tb, d, o)) < m.obb_half_in[k];
    if (faces) {
      const float A = -o * inv, ia = fabsf(inv);
      const float out = m.obb_half_out[k] * ia, in = m.obb_half_in[k] * ia;
      const float slack = 1e-6f * (fabsf(A) + out);  // rounding of the reciprocal, the products and an f64 direction
      nlo[k] = A - out - slack; nhi[k] = A - in + slack;
      flo[k] = A + in - slack; fhi[k] = A + out + slack;
      dk[k] = d;
    }
  }
  if (inside) return 2;
  // slack: 1e-5 relative on both ends of the interval (the slabs are already inflated by 1e-5 x scene magnitude)
  if (!(tn - 1e-5f * fabsf(tn) <= tf + 1e-5f * fabsf(tf))) return 0;
  if (faces) {
    // entry axis: the one whose near interval lies wholly above the other two (at most one can); exit axis likewise below
    const bool e0 = nlo[0] > fmaxf(nhi[1], nhi[2]), e1 = nlo[1] > fmaxf(nhi[0], nhi[2]), e2 = nlo[2] > fmaxf(nhi[0], nhi[1]);
    const bool x0 = fhi[0] < fminf(flo[1], flo[2]), x1 = fhi[1] < fminf(flo[0], flo[2]), x2 = fhi[2] < fminf(flo[0], flo[1]);
    const float nhi_in = e0 ? nhi[0] : (e1 ? nhi[1] : nhi[2]), d_in = e0 ? dk[0] : (e1 ? dk[1] : dk[2]);
    const float flo_out = x0 ? flo[0] : (x1 ? flo[1] : flo[2]), d_out = x0 ? dk[0] : (x1 ? dk[1] : dk[2]);
    // the entry point strictly inside every far plane, the exit point strictly beyond every near plane (for the named
    // axis itself that only asks for a box thicker than the margin)
    const bool sure = (e0 || e1 || e2) && (x0 || x1 || x2) && nhi_in < fminf(flo[0], fminf(flo[1], flo[2])) &&
                      fmaxf(nhi[0], fmaxf(nhi[1], nhi[2])) < flo_out && fabsf(d_in) > 1e-6f && fabsf(d_out) > 1e-6f;
    if (sure) {
      face_in = (e0 ? 0 : (e1 ? 2 : 4)) + (d_in >= 0.f ? 0 : 1);   // moving along +axis enters through the -axis face
      face_out = (x0 ? 0 : (x1 ? 2 : 4)) + (d_out >= 0.f ? 1 : 0);
      return 3;
    }
  }
  return 1;
}

#ifndef RTB_MEDIUM_FACES
#define RTB_MEDIUM_FACES 1
#endif
// BOXSCAN: compile the single-scan path for quad-only boundaries in.  The wavefront shade kernel instantiates
// both and picks per scene: the extra code costs the sphere-media headline scene (c4) 0.9 % through register
// allocation alone, and gains the box-media scene (c3) 32 %.
// face_in / face_out: the entry and exit face medium_obb named for this line (>= 0), -1 = not asked yet, -2 = asked, not certain
template <bool BOXSCAN = true, bool GENERIC = true>
RTB_DEV bool medium_interval(const DScene& S, const DMedium& m, const Ray& r, double& t1, double& t2, int face_in = -1, int face_out = -1) {
  if (!medium_line_cull(m, r)) return false;
  if (BOXSCAN && GENERIC && RTB_MEDIUM_FACES && (m.cls_fast & 0x400)) {
    if (face_in == -1 && medium_obb(m, r, 0.f, 0.f, true, face_in, face_out) != 3) face_in = -2;
    if (face_in >= 0) {
      // the two faces the line certainly crosses: probe 1 (UNIVERSE) returns the entry t, probe 2 over (t1 + 1e-4, inf)
      // the exit t when it lies in that interval (constant_medium.rs:46-55) -- two plane distances instead of six quad tests
      const double ta = quad_t_only(S.prims + (size_t)(m.first_prim + face_in) * PRIM_D2, r);
      const double tb = quad_t_only(S.prims + (size_t)(m.first_prim + face_out) * PRIM_D2, r);
      if (tb >= ta + 0.0001 && tb < RTB_INF && ta > -RTB_INF) { t1 = ta; t2 = tb; return true; }
    }
  }
  if (m.cls_fast & 0x100) {
    // boundary = one static sphere: both probes of constant_medium.rs:46-55 evaluate the SAME two roots
    // (Sphere::hit, object.rs:146-166), so compute them once -- bit-identical to two calls:
    //   probe 1 over UNIVERSE returns root1 (when finite); probe 2 over (root1 + 1e-4, INF) rejects
    //   root1 and returns root2 when it lies in the interval.
    const double2* P = S.prims + (size_t)m.first_prim * PRIM_D2;
    const double2 c01 = RTB_LDG(P + 0), c2r = RTB_LDG(P + 1);
    const double ocx = dsub(r.ox, c01.x), ocy = dsub(r.oy, c01.y), ocz = dsub(r.oz, c2r.x);
    const double a = ddot(r.dx, r.dy, r.dz, r.dx, r.dy, r.dz);
    const double half_b = ddot(ocx, ocy, ocz, r.dx, r.dy, r.dz);
    const double c = dsub(ddot(ocx, ocy, ocz, ocx, ocy, ocz), dmul(c2r.y, c2r.y));
    const double disc = dsub(dmul(half_b, half_b), dmul(a, c));
    if (disc < 0.) return false;
    const double sqrtd = sqrt(disc);
    const double root1 = dsub(-half_b, sqrtd) / a, root2 = dsub(sqrtd, half_b) / a;
    if (-RTB_INF < root1 && root1 < RTB_INF) {
      t1 = root1;
      t2 = root2;
      return (t1 + 0.0001 < root2) && (root2 < RTB_INF);
    }
    // degenerate (zero direction ...): fall through to the generic probes
  }
  // (a scene whose media are all single static spheres never gets past this point with a finite root1: the
  //  generic probes would return +inf for the degenerate rays too, so their code can be compiled out)
  if (!GENERIC) return false;
  if (BOXSCAN && (m.cls_fast & 0x200)) {
    // boundary = quads only (a make_box): a quad's plane distance and inside test do not depend on the probe
    // interval, so ONE scan serves both probes of constant_medium.rs:46-55.  Probe 1 returns the smallest
    // hit distance ta; probe 2 the smallest one >= ta + 1e-4 (Interval::contains is closed), which is the
    // second smallest tb unless two hits lie within 1e-4 of each other (an edge or corner of the box) --
    // then the second probe is run as written.  Same values bit for bit, half the f64 quad tests.
    double ta = RTB_INF, tb = RTB_INF;
    for (int i = 0; i < m.n_prims; i++) {
      double t, a, b;
      if (quad_test(S.prims + (size_t)(m.first_prim + i) * PRIM_D2, r, -RTB_INF, tb, t, a, b)) {
        if (t < ta) { tb = ta; ta = t; }
        else if (t < tb) tb = t;
      }
    }
    t1 = ta;
    if (!(t1 < RTB_INF)) return false;
    if (tb >= t1 + 0.0001) { t2 = tb; return t2 < RTB_INF; }
    t2 = boundary_probe(S, m, r, t1 + 0.0001);
    return t2 < RTB_INF;
  }
  t1 = boundary_probe(S, m, r, -RTB_INF);          // boundary.hit(r, UNIVERSE)        :46
  if (!(t1 < RTB_INF)) return false;
  t2 = boundary_probe(S, m, r, t1 + 0.0001);       // boundary.hit(r, (t1+1e-4, INF))  :49-55
  return t2 < RTB_INF;
}

// (Measured and dropped: for a sphere boundary, bounding the exit distance by |o - c| +- r in fp32 -- to skip the two
//  roots for escaping rays of the book-2 fog -- left the c4 shade stage at 174.9 vs 174.2 ms; the same bound in
//  medium_precheck cost 6 ms.)
// returns the event parameter t (or +inf) for medium `mi`, given the closest surface so far
template <bool BOXSCAN = true, bool GENERIC = true>
RTB_DEV double medium_event(const DScene& S, const DMedium& m, const Ray& r, double tmin, double tmax, float U) {
  const float log_u = logf(U);  // U = 0 -> -inf -> hit_distance +inf: no event
  {  // the shortcut below, first in fp32 with a wide margin (most rays leave here)
    const float len32 = fast_sqrt((float)r.dx * (float)r.dx + (float)r.dy * (float)r.dy + (float)r.dz * (float)r.dz);
    const float span32 = ((float)tmax - (float)tmin) * len32;
    const float hd32 = (float)m.neg_inv_density * log_u;
    if (hd32 > fminf(span32, m.diag) * 1.0001f + 1e-6f) return RTB_INF;
  }
  const double ray_length = sqrt(r.dx * r.dx + r.dy * r.dy + r.dz * r.dz);
  const double hit_distance = m.neg_inv_density * (double)log_u;
  // Exact shortcut: the event needs hit_distance <= (t2' - t1') * |d|, and that product can exceed
  // neither the clamped ray span (tmax - tmin) * |d| nor the longest chord of the boundary (its box
  // diagonal).  A free flight beyond both (with slack for rounding) cannot scatter: skip the two
  // boundary probes.  For the thin book-2 fog (1/rho = 1e4) this removes ~9 of 10 probes.
  {
    const double span = (tmax - tmin) * ray_length;
    const double bound = fmin(span, (double)m.diag);
    if (hit_distance > bound * (1. + 1e-9) + 1e-12) return RTB_INF;
  }
  double t1, t2;
  bool inside = false;
  if ((m.cls_fast & 0x100) && tmax < RTB_INF) {
    // Boundary = one static sphere and the clamped segment [tmin, tmax] lies wholly inside it (both end points do,
    // by a wide fp32 margin; a ball is convex): then root1 < 0 < tmin and root2 > tmax, both probes of
    // constant_medium.rs:46-55 succeed, and the clamps below turn the interval into [tmin, tmax] whatever the
    // roots are -- the same value without evaluating them.  (The book-2 fog of radius 5000 contains every segment
    // that ends on a surface: two f64 divisions and a square root per probe, gone.)
    const float ax = (float)r.ox - m.sphere[0], ay = (float)r.oy - m.sphere[1], az = (float)r.oz - m.sphere[2];
    const float dx = (float)r.dx, dy = (float)r.dy, dz = (float)r.dz, ta = (float)tmin, tb = (float)tmax;
    const float pa = fmaf(ta, dx, ax), qa = fmaf(ta, dy, ay), ra = fmaf(ta, dz, az);
    const float pb = fmaf(tb, dx, ax), qb = fmaf(tb, dy, ay), rb = fmaf(tb, dz, az);
    const float da = fmaf(pa, pa, fmaf(qa, qa, ra * ra)), db = fmaf(pb, pb, fmaf(qb, qb, rb * rb));
    inside = da < m.sphere[3] && db < m.sphere[3];  // sphere[3] = r^2 shrunk by 1e-3 relative (flatten.cpp)
  }
  int face_in = -1, face_out = -1;
  if (BOXSCAN && GENERIC && (m.cls_fast & 0x400)) {
    const int where = medium_obb(m, r, (float)tmin, tmax < RTB_INF ? (float)tmax : 3.0e38f, RTB_MEDIUM_FACES != 0, face_in, face_out);
    if (where == 0) return RTB_INF;
    inside = where == 2 && tmax < RTB_INF;
    if (where != 3) face_in = -2;
  }
  if (inside) { t1 = -RTB_INF; t2 = RTB_INF; }
  else if (!medium_interval<BOXSCAN, GENERIC>(S, m, r, t1, t2, face_in, face_out)) return RTB_INF;
  if (t1 < tmin) t1 = tmin;   // :58-60
  if (t2 > tmax) t2 = tmax;   // :61-63
  if (t1 >= t2) return RTB_INF;
  if (t1 < 0.) t1 = 0.;       // :69-71
  const double distance_inside_boundary = (t2 - t1) * ray_length;
  if (hit_distance > distance_inside_boundary) return RTB_INF;
  return t1 + hit_distance / ray_length;
}

// Wavefront front end of medium_event: the ray arrives as stored in the queue (f64 origin, fp32
// direction).  Two conservative fp32 rejections -- the free-flight shortcut (fast log, wide margin)
// and the boundary-box line cull -- run before anything is widened to f64; both only ever skip
// events that medium_event itself would reject, so the two pipelines take identical decisions.
// medium_precheck: false = certainly no event in (tmin, tmax); true = evaluate medium_event.
RTB_DEV bool medium_precheck(const DMedium& m, double ox, double oy, double oz, float dx, float dy, float dz, double tmin,
                             double tmax, float U) {
#if defined(__CUDACC__)
  const float fast_log = __logf(U);
#else
  const float fast_log = logf(U);
#endif
  const float len32 = fast_sqrt(dx * dx + dy * dy + dz * dz);
  const float span32 = ((float)tmax - (float)tmin) * len32;
  const float nid = (float)m.neg_inv_density;
  // __logf: abs error <= 2^-21.4 on (0.5, 2), else 2 ulp -> slack nid * 1e-6 plus 0.1 % relative
  if (nid * fast_log > fminf(span32, m.diag) * 1.001f - nid * 1e-6f + 1e-5f) return false;
  const float fx = (float)ox, fy = (float)oy, fz = (float)oz;
  const float idx = fast_rcp(dx), idy = fast_rcp(dy), idz = fast_rcp(dz);
  const float a0 = (m.lo[0] - fx) * idx, a1 = (m.hi[0] - fx) * idx;
  const float b0 = (m.lo[1] - fy) * idy, b1 = (m.hi[1] - fy) * idy;
  const float c0 = (m.lo[2] - fz) * idz, c1 = (m.hi[2] - fz) * idz;
  const float tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fminf(c0, c1));
  const float tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fmaxf(c0, c1));
  return tn <= fmaf(fabsf(tf), 2e-6f, tf) + 1e-30f;
}

template <bool BOXSCAN = true, bool GENERIC = true>
RTB_DEV double medium_event_lazy(const DScene& S, const DMedium& m, const Ray& r, double tmin, double tmax, float U) {
  // (a primary ray's f64 direction is rounded for the precheck only: 6e-8 relative against margins of 1e-3 and 2e-6)
  if (!medium_precheck(m, r.ox, r.oy, r.oz, (float)r.dx, (float)r.dy, (float)r.dz, tmin, tmax, U)) return RTB_INF;
  return medium_event<BOXSCAN, GENERIC>(S, m, r, tmin, tmax, U);
}

// one sample's radiance as the fixed-point addend of the accumulation buffer (device_scene.h ACCUM_SCALE);
// the conversion saturates, a finite sample can at worst pin its pixel
RTB_DEV unsigned long long accum_fixed(float L) {
#if defined(__CUDACC__)
  return (unsigned long long)__double2ll_rn((double)L * ACCUM_SCALE);
#else
  const double v = (double)L * ACCUM_SCALE;
  return (unsigned long long)(long long)(v >= 9.2e18 ? 9.2e18 : (v <= -9.2e18 ? -9.2e18 : nearbyint(v)));
#endif
}

// ------------------------------------------------------------------------------------------------
// textures  (src/texture.rs:18-131, src/perlin.rs:30-96, src/rt_image.rs:37-46)
// ------------------------------------------------------------------------------------------------
RTB_DEV float perlin_noise(const DScene& S, int table, double px, double py, double pz) {  // perlin.rs:30-54,74-96
  const double fx = floor(px), fy = floor(py), fz = floor(pz);
  const float u = (float)(px - fx), v = (float)(py - fy), w = (float)(pz - fz);
  // `as i32` saturates; |p| stays far below 2^31 for every octave of the scenes in scope
  const int i = (int)fmax(-2147483648.0, fmin(2147483647.0, fx));
  const int j = (int)fmax(-2147483648.0, fmin(2147483647.0, fy));
  const int k = (int)fmax(-2147483648.0, fmin(2147483647.0, fz));
  const float uu = u * u * (3.f - 2.f * u), vv = v * v * (3.f - 2.f * v), ww = w * w * (3.f - 2.f * w);
  const uint8_t* perm = S.perlin_perm + 768 * table;
  const float4* vec = S.perlin_vec + 256 * table;
  float accum = 0.f;
#pragma unroll
  for (int di = 0; di < 2; di++)
#pragma unroll
    for (int dj = 0; dj < 2; dj++)
#pragma unroll
      for (int dk = 0; dk < 2; dk++) {
        const int idx = RTB_LDG(perm + ((i + di) & 255)) ^ RTB_LDG(perm + 256 + ((j + dj) & 255)) ^
                        RTB_LDG(perm + 512 + ((k + dk) & 255));
        const float4 c = RTB_LDG(vec + idx);
        const float wx = di ? uu : 1.f - uu, wy = dj ? vv : 1.f - vv, wz = dk ? ww : 1.f - ww;
        accum += wx * wy * wz * (c.x * (u - di) + c.y * (v - dj) + c.z * (w - dk));
      }
  return accum;
}

RTB_DEV float perlin_turb(const DScene& S, int table, double px, double py, double pz) {  // perlin.rs:56-72
  float accum = 0.f, weight = 1.f;
  for (int o = 0; o < 7; o++) {  // (unrolled by the compiler; "#pragma unroll 1" saves 9 KB of code and measured no faster)
    accum += weight * perlin_noise(S, table, px, py, pz);
    weight *= 0.5f;
    px *= 2.; py *= 2.; pz *= 2.;
  }
  return fabsf(accum);
}

// TEXTURES = false: every texture of the scene is a solid colour, the checker / image / noise code is compiled out
template <bool TEXTURES = true>
RTB_DEV V3 texture_value(const DScene& S, const DTexture* textures, int ti, float u, float v, double px, double py, double pz) {
  if (!TEXTURES) {
    const DTexture& t = textures[ti];
    return v3(t.color[0], t.color[1], t.color[2]);
  }
  for (int guard = 0; guard < 16; guard++) {
    const DTexture& t = textures[ti];
    if (t.kind == TEX_SOLID) return v3(t.color[0], t.color[1], t.color[2]);  // texture.rs:44-46
    if (t.kind == TEX_CHECKER) {  // texture.rs:71-81 (Q19): floor in f64, Rust `%` keeps the sign
      const int x = (int)floor(t.scale * px), y = (int)floor(t.scale * py), z = (int)floor(t.scale * pz);
      const int sum = (int)((unsigned)x + (unsigned)y + (unsigned)z);  // wrapping add, like a release build
      ti = ((sum % 2) == 0) ? t.a : t.b;
      continue;
    }
    if (t.kind == TEX_IMAGE) {  // texture.rs:95-107, rt_image.rs:37-46 (Q20)
      if (t.height <= 0) return v3(0.f, 1.f, 1.f);
      const float uc = __saturatef(u), vc = __saturatef(v);
      unsigned i = (unsigned)(uc * (float)t.width), j = (unsigned)(vc * (float)t.height);
      if (i > (unsigned)t.width - 1u) i = (unsigned)t.width - 1u;
      unsigned y = (unsigned)t.height - j - 1u;  // wraps when j == height, then clamps
      if (y > (unsigned)t.height - 1u) y = (unsigned)t.height - 1u;
      const uint8_t* p = S.texels + t.a + 3 * ((size_t)y * t.width + i);
      const float s = 1.0f / 255.0f;
      return v3(RTB_LDG(p) * s, RTB_LDG(p + 1) * s, RTB_LDG(p + 2) * s);
    }
    // TEX_NOISE  texture.rs:127-130 (Q21)
    const double sx = t.scale * px, sy = t.scale * py, sz = t.scale * pz;
    const float g = 0.5f * (1.f + sinf((float)sz + 10.f * perlin_turb(S, t.a, sx, sy, sz)));
    return v3(g, g, g);
  }
  return v3(0.f, 0.f, 0.f);
}
RTB_DEV V3 texture_value(const DScene& S, int ti, float u, float v, double px, double py, double pz) {
  return texture_value<true>(S, S.textures, ti, u, v, px, py, pz);
}

// ------------------------------------------------------------------------------------------------
// sampling maps and pdfs (src/onb.rs, src/pdf.rs, src/vec3.rs:184-250) -- rejection loops are
// replaced by direct maps with a fixed draw count (same distributions)
// ------------------------------------------------------------------------------------------------
struct Onb { V3 u, v, w; };
RTB_DEV Onb onb_from_w(V3 w_in) {  // onb.rs:32-47
  Onb o;
  o.w = normalize(w_in);
  const V3 a = fabsf(o.w.x) > 0.9f ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
  o.v = normalize(cross(o.w, a));
  o.u = cross(o.w, o.v);
  return o;
}
RTB_DEV V3 onb_local(const Onb& o, float a, float b, float c) { return a * o.u + (b * o.v + c * o.w); }  // onb.rs:24-30

RTB_DEV V3 random_cosine_direction(float r1, float r2) {  // vec3.rs:240-250
  float s, c;
  sincospif(2.f * r1, &s, &c);
  const float sr = fast_sqrt(r2);
  return v3(c * sr, s * sr, fast_sqrt(1.f - r2));
}
RTB_DEV V3 random_unit_vector(float r1, float r2) {  // uniform sphere; vec3.rs:215-217 by direct map
  const float z = 1.f - 2.f * r1;
  const float rr = fast_sqrt(fmaxf(0.f, 1.f - z * z));
  float s, c;
  sincospif(2.f * r2, &s, &c);
  return v3(rr * c, rr * s, z);
}

// get_sphere_uv  src/object.rs:114-120 (Q8) in f64 (parity harness and IMAGE textures)
RTB_DEV void sphere_uv(double nx, double ny, double nz, double& u, double& v) {
  const double theta = acos(-ny);
  const double phi = atan2(-nz, nx) + PI_D;
  const double inv_pi = 1.0 / PI_D;
  u = phi * inv_pi * 0.5;
  v = theta * inv_pi;
}

// ------------------------------------------------------------------------------------------------
// light list: HittableList::{pdf_value, random} over Quad/Sphere::{pdf_value, random}
// (src/hittable.rs:115-129, src/object.rs:122-132, 190-212, 492-506; Q9-Q12)
// ------------------------------------------------------------------------------------------------
RTB_DEV double light_pdf_one(const DLight& L, const Ray& probe) {
  const double2* P = reinterpret_cast<const double2*>(L.prim);
  if (L.kind == LIGHT_QUAD) {  // Quad::pdf_value :492-501
    double t, a, b;
    if (!quad_test(P, probe, 0.001, RTB_INF, t, a, b)) return 0.;
    const double len2 = probe.dx * probe.dx + probe.dy * probe.dy + probe.dz * probe.dz;
    const double distance_squared = t * t * len2;
    const double cosine = fabs((probe.dx * L.prim[0] + probe.dy * L.prim[1] + probe.dz * L.prim[2]) / sqrt(len2));
    return distance_squared / (cosine * L.area);
  }
  if (L.kind == LIGHT_SPHERE) {  // Sphere::pdf_value :190-202 (time 0, motion ignored)
    double t;
    if (!sphere_test(P, 0, probe, 0., 0.001, RTB_INF, t)) return 0.;
    const double ex = L.prim[0] - probe.ox, ey = L.prim[1] - probe.oy, ez = L.prim[2] - probe.oz;
    const double cos_theta_max = sqrt(1. - L.prim[3] * L.prim[3] / (ex * ex + ey * ey + ez * ez));
    return 1. / (2. * PI_D * (1. - cos_theta_max));
  }
  return 0.;  // Hittable default  src/hittable.rs:46-48
}

RTB_DEV float lights_pdf_value(const DScene& S, double ox, double oy, double oz, V3 dir) {
  Ray probe;
  probe.ox = ox; probe.oy = oy; probe.oz = oz;
  probe.dx = (double)dir.x; probe.dy = (double)dir.y; probe.dz = (double)dir.z;
  probe.time = 0.;  // Ray::new  src/ray.rs:20-26
  double sum = 0.;
  for (int i = 0; i < S.n_lights; i++) sum += light_pdf_one(S.lights[i], probe);
  return (float)(sum * (1. / (double)S.n_lights));
}

RTB_DEV V3 lights_random(const DScene& S, double ox, double oy, double oz, float upick, float r1, float r2) {
  int pick = (int)(upick * (float)S.n_lights);  // random_int(0, n-1)  src/hittable.rs:127-128
  if (pick > S.n_lights - 1) pick = S.n_lights - 1;
  const DLight& L = S.lights[pick];
  if (L.kind == LIGHT_QUAD) {  // Quad::random :503-506
    const double a = (double)r1, b = (double)r2;
    return v3((float)(L.q[0] + a * L.u[0] + b * L.v[0] - ox), (float)(L.q[1] + a * L.u[1] + b * L.v[1] - oy),
              (float)(L.q[2] + a * L.u[2] + b * L.v[2] - oz));
  }
  if (L.kind == LIGHT_SPHERE) {  // Sphere::random :204-212, random_to_sphere :122-132
    const double ex = L.prim[0] - ox, ey = L.prim[1] - oy, ez = L.prim[2] - oz;
    const double dist2 = ex * ex + ey * ey + ez * ez;
    const Onb uvw = onb_from_w(v3((float)ex, (float)ey, (float)ez));
    const float z = 1.f + r2 * ((float)sqrt(1. - L.prim[3] * L.prim[3] / dist2) - 1.f);
    float s, c;
    sincospif(2.f * r1, &s, &c);
    const float sq = fast_sqrt(1.f - z * z);
    return onb_local(uvw, c * sq, s * sq, z);
  }
  return v3(1.f, 0.f, 0.f);  // Hittable default  src/hittable.rs:50-52
}

// full hit record of the winner, f64 (parity harness): Sphere::hit :168-183 / Quad::hit :477-489 and
// the point/normal mapping of Translate/RotateY (already in world space here because instances are baked)
RTB_DEV void complete_hit(const DScene& S, const Ray& r, const Hit& best, RtbHit& out) {
  out.prim = -1; out.front_face = 0; out.material = -1; out.reserved = 0;
  out.t = RTB_INF;
  out.p[0] = out.p[1] = out.p[2] = 0.; out.normal[0] = out.normal[1] = out.normal[2] = 0.;
  out.u = out.v = 0.;
  if (best.prim < 0) return;
  RTB_ASSERT(best.prim < S.n_prims);
  const int4 info = RTB_LDG(S.prim_info + best.prim);
  RTB_ASSERT(info.z >= 0 && info.z < 64);
  const double2* P = S.prims + (size_t)best.prim * PRIM_D2;
  const double t = best.t;
  const double px = dadd(r.ox, dmul(t, r.dx)), py = dadd(r.oy, dmul(t, r.dy)), pz = dadd(r.oz, dmul(t, r.dz));  // Ray::at
  double nx, ny, nz, u, v;
  if ((info.x & 0xFF) == PRIM_QUAD) {
    nx = P[0].x; ny = P[0].y; nz = P[1].x;
    u = best.a; v = best.b;
  } else {
    double cx = P[0].x, cy = P[0].y, cz = P[1].x;
    if (info.x & PRIM_FLAG_MOVING) {
      cx = dadd(cx, dmul(r.time, P[2].x)); cy = dadd(cy, dmul(r.time, P[2].y)); cz = dadd(cz, dmul(r.time, P[3].x));
    }
    const double rad = P[1].y;  // (p - center) / radius is a true division (src/vec3.rs:145-151)
    nx = dsub(px, cx) / rad; ny = dsub(py, cy) / rad; nz = dsub(pz, cz) / rad;
    const double2 cs = S.xforms[info.z];
    sphere_uv(cs.x * nx - cs.y * nz, ny, cs.y * nx + cs.x * nz, u, v);
  }
  const bool front = ddot(r.dx, r.dy, r.dz, nx, ny, nz) < 0.;
  if (!front) { nx = -nx; ny = -ny; nz = -nz; }
  out.prim = info.w; out.front_face = front ? 1 : 0; out.material = info.y;
  out.t = t; out.p[0] = px; out.p[1] = py; out.p[2] = pz;
  out.normal[0] = nx; out.normal[1] = ny; out.normal[2] = nz;
  out.u = u; out.v = v;
}

// ------------------------------------------------------------------------------------------------
// one path segment: extend (closest event) + shade (emit / scatter / terminate)
// ------------------------------------------------------------------------------------------------
struct PathState {
  Ray ray;
  float bx, by, bz;  // throughput (product of attenuation * scattering_pdf / pdf)
  uint32_t pixel, sample, bounce;
};

// get_ray  src/render.rs:218-249 (draw order: jitter x, jitter y, [disk], time)
RTB_DEV void generate_primary(const DScene& S, uint32_t pixel, uint32_t sample, PathState& ps) {
  const DCamera& cam = S.cam;
  const int x = (int)(pixel % (uint32_t)cam.width), y = (int)(pixel / (uint32_t)cam.width);
  const int s_j = (int)(sample / (uint32_t)cam.sqrt_spp), s_i = (int)(sample % (uint32_t)cam.sqrt_spp);
  const Rand4 u = rand4(S, pixel, sample, PRIMARY_BOUNCE, 0);
  const double px = (double)x + (-0.5 + cam.recip_sqrt_spp * ((double)s_i + (double)u.x));  // :246
  const double py = (double)y + (-0.5 + cam.recip_sqrt_spp * ((double)s_j + (double)u.y));  // :247
  const double sx = cam.pixel00[0] + px * cam.du[0] + py * cam.dv[0];
  const double sy = cam.pixel00[1] + px * cam.du[1] + py * cam.dv[1];
  const double sz = cam.pixel00[2] + px * cam.du[2] + py * cam.dv[2];
  double ox = cam.center[0], oy = cam.center[1], oz = cam.center[2];
  if (cam.defocus) {  // :226-230, 238-241; unit disk by the polar map instead of rejection
    const Rand4 d = rand4(S, pixel, sample, PRIMARY_BOUNCE, 1);
    const float rr = fast_sqrt(d.x);
    float s, c;
    sincospif(2.f * d.y, &s, &c);
    const double dxk = (double)(rr * c), dyk = (double)(rr * s);
    ox += dxk * cam.disk_u[0] + dyk * cam.disk_v[0];
    oy += dxk * cam.disk_u[1] + dyk * cam.disk_v[1];
    oz += dxk * cam.disk_u[2] + dyk * cam.disk_v[2];
  }
  ps.ray.ox = ox; ps.ray.oy = oy; ps.ray.oz = oz;
  ps.ray.dx = sx - ox; ps.ray.dy = sy - oy; ps.ray.dz = sz - oz;
  ps.ray.time = (double)u.z;  // :233
  ps.bx = ps.by = ps.bz = 1.f;
  ps.pixel = pixel; ps.sample = sample; ps.bounce = 0;
}

struct Event {
  double t;   // +inf: miss
  double a, b;
  int prim;   // >= 0 surface primitive (BVH order); -1 none
  int medium; // >= 0: the event is a scatter inside this medium
  int have_ab;  // quad planar coordinates a, b are valid (else shade recomputes them when a texture needs uv)
  int info_x;   // prim_info[prim].x when the caller already has it (wavefront hit records carry it), else 0
};

// world.hit(r, Interval{0.0001, INF})  src/render.rs:264-270
template <bool STATS>
RTB_DEV void extend(const DScene& S, const PathState& ps, Event& ev, DStats* st) {
  Hit best;
  hit_reset(best);
  if (S.n_surface_prims > 0) closest_surface<STATS>(S, ps.ray, 0.0001, best, st);
  ev.t = best.t; ev.a = best.a; ev.b = best.b; ev.prim = best.prim; ev.medium = -1; ev.have_ab = 1; ev.info_x = 0;
  if (S.n_media > 0) {
    Rand4 u;
    for (int mi = 0; mi < S.n_media; mi++) {
      if ((mi & 3) == 0) u = rand4(S, ps.pixel, ps.sample, ps.bounce, 1u + (uint32_t)(mi >> 2));
      const float U = (mi & 3) == 0 ? u.x : ((mi & 3) == 1 ? u.y : ((mi & 3) == 2 ? u.z : u.w));
      if (STATS) st->medium_probes++;
      const double tm = medium_event(S, S.media[mi], ps.ray, 0.0001, ev.t, U);
      if (tm < ev.t) { ev.t = tm; ev.medium = mi; ev.prim = -1; }
    }
  }
}

// Dielectric::scatter  material.rs:167-191 (Q15): Schlick reflectance :156-163, reflect / refract  vec3.rs:219-229.
// `n` is the face normal (against the ray), `u` the uniform draw (consumed only when refraction is possible).
RTB_DEV V3 dielectric_direction(V3 d, V3 n, bool front, float ir, float u) {
  const float ratio = front ? fast_rcp(ir) : ir;
  const V3 ud = normalize(d);
  const float cos_theta = fminf(dot(-ud, n), 1.f);
  const float sin_theta = fast_sqrt(1.f - cos_theta * cos_theta);
  bool reflect_it = ratio * sin_theta > 1.f;
  if (!reflect_it) {
    float r0 = fast_div(1.f - ratio, 1.f + ratio);
    r0 = r0 * r0;
    const float x = 1.f - cos_theta, x2 = x * x;
    reflect_it = (r0 + (1.f - r0) * (x2 * x2 * x)) > u;
  }
  if (reflect_it) return ud - (2.f * dot(ud, n)) * n;
  const V3 perp = ratio * (ud + cos_theta * n);
  const V3 par = -fast_sqrt(fabsf(1.f - dot(perp, perp))) * n;
  return perp + par;
}

// material response at the event; returns false when the path ends (contribution added to L)
// LIGHTS = false compiles the light-list sampler (HittablePDF: f64 probes per light) out: the wavefront shade
// kernel instantiates it for scenes whose light list is empty (what render_par passes, F2).
// QUAD_UV / SPHERE_UV = false compile the (u, v) evaluation of that primitive kind out (no material reads it).
template <bool LIGHTS = true, bool QUAD_UV = true, bool SPHERE_UV = true, bool TEXTURES = true>
RTB_DEV bool shade(const DScene& S, const Tables& T, PathState& ps, const Event& ev, float& Lr, float& Lg, float& Lb,
                   DStats* st, bool stats) {
  const Ray& r = ps.ray;
  if (!(ev.t < RTB_INF)) {  // miss: cam.background  src/render.rs:298-309
    float er = S.cam.background[0], eg = S.cam.background[1], eb = S.cam.background[2];
    if (S.n_suns > 0) {  // `cam.background + sun_light` (the term HEAD comments out; RTB_FLAG_SUN_LIGHT): Sun::_hit  object.rs:232-239
      const V3 ud = normalize(v3((float)r.dx, (float)r.dy, (float)r.dz));
      for (int k = 0; k < S.n_suns; k++)
        if (dot(ud, v3(S.suns[k].dir[0], S.suns[k].dir[1], S.suns[k].dir[2])) > S.suns[k].limit) {
          er += S.suns[k].albedo[0]; eg += S.suns[k].albedo[1]; eb += S.suns[k].albedo[2];
        }
    }
    Lr += ps.bx * er; Lg += ps.by * eg; Lb += ps.bz * eb;
    return false;
  }
  const double px = dadd(r.ox, dmul(ev.t, r.dx)), py = dadd(r.oy, dmul(ev.t, r.dy)), pz = dadd(r.oz, dmul(ev.t, r.dz));  // Ray::at
  V3 n;
  bool front = true;
  float tu = 0.f, tv = 0.f;
  int mat_id;
  if (ev.medium >= 0) {  // constant_medium.rs:82-90: arbitrary normal, front_face = true, u = v = 0
    n = v3(1.f, 0.f, 0.f);
    mat_id = T.media[ev.medium].material;
  } else {
    // kind | flags | class | material: from the hit record when it carries them, else one gather
    int info_x = ev.info_x;
    if (info_x == 0) info_x = RTB_LDG(S.prim_info + ev.prim).x;
    const double2* P = S.prims + (size_t)ev.prim * PRIM_D2;
    mat_id = (info_x >> PRIM_MAT_SHIFT) & 0xFFF;
    mat_id = mat_id ? mat_id - 1 : RTB_LDG(S.prim_info + ev.prim).y;
    const int needs_uv = T.materials[mat_id].needs_uv;
    if ((info_x & 0xFF) == PRIM_QUAD) {
      const double2 n01 = RTB_LDG(P + 0), n2d = RTB_LDG(P + 1);
      front = ddot(r.dx, r.dy, r.dz, n01.x, n01.y, n2d.x) < 0.;  // set_face_normal  hittable.rs:22-37
      n = v3((float)n01.x, (float)n01.y, (float)n2d.x);
      if (QUAD_UV && needs_uv) {
        double ta = ev.a, tb = ev.b;
        if (!ev.have_ab) {
          ta = tb = 0.;  // wavefront hit records carry only (t, id): re-evaluate Quad::hit for alpha/beta
          double tt;
          quad_test(P, r, -RTB_INF, RTB_INF, tt, ta, tb);
        }
        tu = (float)ta; tv = (float)tb;
      }
    } else {
      const double2 c01 = RTB_LDG(P + 0), c2r = RTB_LDG(P + 1);
      double cx = c01.x, cy = c01.y, cz = c2r.x;
      if (info_x & PRIM_FLAG_MOVING) {
        const double2 v01 = RTB_LDG(P + 2), v2 = RTB_LDG(P + 3);
        cx += r.time * v01.x; cy += r.time * v01.y; cz += r.time * v2.x;
      }
      // outward normal (p - c) / r  (object.rs:169): the difference in f64 (p and c are large, their
      // difference is not), the scaling in fp32 -- an f64 division per sphere hit buys nothing here
      const double ex = px - cx, ey = py - cy, ez = pz - cz;
      const float inv_r = fast_rcp((float)c2r.y);
      n = v3((float)ex * inv_r, (float)ey * inv_r, (float)ez * inv_r);
      front = (r.dx * ex + r.dy * ey + r.dz * ez) * c2r.y < 0.;  // sign of d.n with n = e / r (a negative radius flips it)
      if (SPHERE_UV && needs_uv) {  // uv live in object space: undo the baked rotate_y (transform.rs:85-105); f64 like get_sphere_uv
        const double inv_r64 = 1. / c2r.y;
        const double nx = ex * inv_r64, ny = ey * inv_r64, nz = ez * inv_r64;
        const double2 cs = RTB_LDG(S.xforms + RTB_LDG(S.prim_info + ev.prim).z);
        double u64, v64;
        sphere_uv(cs.x * nx - cs.y * nz, ny, cs.y * nx + cs.x * nz, u64, v64);
        tu = (float)u64; tv = (float)v64;
      }
    }
    if (!front) n = -n;
  }
  const DMaterial& m = T.materials[mat_id];
  if (m.kind == MAT_DIFFUSE_LIGHT) {  // emitted: front face only; never scatters (Q16)  material.rs:210-221
    if (front) {
      const V3 e = texture_value<TEXTURES>(S, T.textures, m.texture, tu, tv, px, py, pz);
      Lr += ps.bx * e.x; Lg += ps.by * e.y; Lb += ps.bz * e.z;
    }
    return false;
  }
  const Rand4 u = rand4(S, ps.pixel, ps.sample, ps.bounce, 0);
  V3 dir;
  if (m.kind == MAT_METAL) {  // material.rs:125-134 (Q14): always scatters
    const V3 ud = normalize(v3((float)r.dx, (float)r.dy, (float)r.dz));
    const V3 refl = ud - (2.f * dot(ud, n)) * n;
    dir = normalize(refl) + m.param * random_unit_vector(u.z, u.w);
    ps.bx *= m.color[0]; ps.by *= m.color[1]; ps.bz *= m.color[2];
  } else if (m.kind == MAT_DIELECTRIC) {
    dir = dielectric_direction(v3((float)r.dx, (float)r.dy, (float)r.dz), n, front, m.param, u.x);
    ps.bx *= m.color[0]; ps.by *= m.color[1]; ps.bz *= m.color[2];
  } else {
    // PdfPtr arm  src/render.rs:278-293: MixturePDF(HittablePDF(lights), material pdf)  pdf.rs:102-127
    const V3 atten = texture_value<TEXTURES>(S, T.textures, m.texture, tu, tv, px, py, pz);
    const bool lambert = m.kind == MAT_LAMBERTIAN;
    Onb uvw;
    if (lambert) uvw = onb_from_w(n);  // CosinePDF::new  pdf.rs:60-66
    const bool have_lights = LIGHTS && S.n_lights > 0;  // empty list: material pdf alone (F2)
    if (have_lights && u.x < 0.5f) {
      dir = lights_random(S, px, py, pz, u.y, u.z, u.w);
    } else if (lambert) {
      const V3 c = random_cosine_direction(u.z, u.w);
      dir = onb_local(uvw, c.x, c.y, c.z);
    } else {
      dir = random_unit_vector(u.z, u.w);  // SpherePDF::generate  pdf.rs:51-53
    }
    const V3 udir = normalize(dir);
    float mat_pdf, scattering_pdf;
    if (lambert) {
      mat_pdf = fmaxf(0.f, dot(udir, uvw.w) * (1.0f / PI_F));              // CosinePDF::value  pdf.rs:69-73
      const float cos_theta = dot(n, udir);                                 // Lambertian::scattering_pdf :100-109
      scattering_pdf = cos_theta < 0.f ? 0.f : cos_theta * (1.0f / PI_F);
    } else {
      mat_pdf = 1.0f / (4.0f * PI_F);                                       // SpherePDF::value  pdf.rs:47-49
      scattering_pdf = (S.flags & 1u) ? 0.f : 1.0f / (4.0f * PI_F);         // F3 (RTB_FLAG_ISO_PDF_ZERO)
    }
    float pdf_val = mat_pdf;
    if (have_lights) pdf_val = 0.5f * lights_pdf_value(S, px, py, pz, dir) + 0.5f * mat_pdf;  // pdf.rs:116-118
    if (!(S.flags & 2u)) {  // default NaN policy (Q22): zero / non-finite pdf -> the sample contributes nothing
      if (!(pdf_val > 0.f) || !(pdf_val < 3.0e38f)) {
        if (stats) st->nonfinite++;
        return false;
      }
    }
    const float wgt = fast_div(scattering_pdf, pdf_val);
    ps.bx *= atten.x * wgt; ps.by *= atten.y * wgt; ps.bz *= atten.z * wgt;
    if (!(S.flags & 2u) && ps.bx == 0.f && ps.by == 0.f && ps.bz == 0.f) return false;  // dead path: result is 0
  }
  if ((S.flags & 0x200u) && ps.bounce >= 3u) {  // RTB_FLAG_RUSSIAN_ROULETTE (opt-in; the reference has none)
    const float q = fminf(1.f, fmaxf(0.05f, fmaxf(ps.bx, fmaxf(ps.by, ps.bz))));
    if (!(rand4(S, ps.pixel, ps.sample, ps.bounce, 63u).x < q)) return false;
    const float inv_q = fast_rcp(q);
    ps.bx *= inv_q; ps.by *= inv_q; ps.bz *= inv_q;
  }
  ps.ray.ox = px; ps.ray.oy = py; ps.ray.oz = pz;  // Ray::new_timed(rec.p, dir, r.time())
  ps.ray.dx = (double)dir.x; ps.ray.dy = (double)dir.y; ps.ray.dz = (double)dir.z;
  ps.bounce++;
  return ps.bounce < (uint32_t)S.cam.max_depth;  // depth <= 0 -> (0,0,0)  src/render.rs:260-262
}
RTB_DEV bool shade(const DScene& S, PathState& ps, const Event& ev, float& Lr, float& Lg, float& Lb, DStats* st,
                   bool stats) {
  return shade<true, true, true, true>(S, scene_tables(S), ps, ev, Lr, Lg, Lb, st, stats);
}

}  // namespace rtb
