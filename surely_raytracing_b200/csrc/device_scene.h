// device_scene.h -- the flat, device-resident form of a scene (host builder: flatten.cpp,
// consumers: the kernels in kernels.cu).  Layout in HBM (all L2-resident: C4 is < 0.6 MB):
//
//   nodes      float4[4*n_nodes]   BVH2, one 64 B record per inner node holding BOTH children's
//                                  boxes (fp32, padded outward so the fp32 slab test is conservative)
//   nodes4     float4[8*n_nodes4]  the same tree with every other level collapsed (BVH4, 128 B per node:
//                                  lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4] refs[4] pad): half the
//                                  dependent node steps per ray for the latency-bound wavefront extend
//   qnodes     uint4[2*n_nodes]    the same tree in 32 B per node: child boxes as 16-bit cell indices
//                                  of a scene-wide grid, rounded outward (one 32 B sector per visit)
//   prims      double2[8*n_prims]  128 B per primitive in BVH-leaf order, f64, world space (instance
//                                  transforms baked in); surfaces first, then medium-boundary prims
//   pre        DPre[n_prims]       64 B per primitive: the quad fields of the conservative fp32 prefilter
//   prim_info  int4[n_prims]       {kind | flags | class << 16 | (material + 1) << 20, material, xform, canonical id}
//   xforms     double2[n_xforms]   {cos, sin} of the composed rotate_y of an instance chain (uv only)
//   media, materials, textures, texels (u8 RGB), perlin tables, lights: small tagged records.
//
// Precision split (DESIGN.md "Numerics"): BVH culling is fp32 and conservative; every
// accept/reject decision that can change a result (primitive tests, medium intervals, light-pdf
// probes) is f64 like the reference; shading arithmetic is fp32.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) || __has_include(<vector_types.h>)
#include <vector_types.h>
#else
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
#endif

namespace rtb {

enum : int { PRIM_SPHERE = 0, PRIM_QUAD = 1 };
enum : int { PRIM_FLAG_MOVING = 0x100 };
// shading class of the material behind a primitive / medium (bits 16..19 of prim_info.x): the sort key
// of the wavefront shade stage.  Lambertian is split by texture so that the expensive Perlin
// evaluation and the cheap solid colour never share a warp.
enum : int { CLS_MISS = 0, CLS_LIGHT = 1, CLS_LAMBERT_SOLID = 2, CLS_LAMBERT_TEX = 3, CLS_METAL = 4, CLS_DIELECTRIC = 5,
             CLS_ISOTROPIC = 6, CLS_NOISE = 7, NUM_CLASSES = 8 };
constexpr int PRIM_CLASS_SHIFT = 16;
// bits 20..31 of prim_info.x: material index + 1 (0 = does not fit, read prim_info.y), so that the one
// word a hit record carries tells the shade stage kind, class and material without a gather
constexpr int PRIM_MAT_SHIFT = 20;
constexpr int PRIM_MAT_MAX = 4094;
enum : int { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4 };
enum : int { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_IMAGE = 2, TEX_NOISE = 3 };
enum : int { LIGHT_QUAD = 0, LIGHT_SPHERE = 1, LIGHT_OTHER = 2 };

// Features a scene may or may not use; code for an unused one is compiled out of the shade instantiation the
// scene runs (a set bit never changes a result, it only keeps code in).
enum : int { SPEC_MEDIA = 1,        // constant media present
             SPEC_BOXSCAN = 2,      // a medium bounded by quads only (single-scan boundary interval)
             SPEC_LIGHTS = 4,       // non-empty light list (HittablePDF sampling)
             SPEC_GENERIC_MEDIA = 8,  // a medium whose boundary is not one static sphere (generic boundary probes)
             SPEC_QUAD_UV = 16,     // a quad whose material reads (u, v)
             SPEC_SPHERE_UV = 32,   // a sphere whose material reads (u, v)
             SPEC_TEXTURES = 64,    // a texture that is not a solid colour (checker / image / noise evaluation)
             SPEC_MULTI_LEAF = 128, // BVH leaves of several primitives (RTB_FLAG_BVH_LEAF4): candidates are whole leaves
             SPEC_ALL = 255 };

// Accumulation buffer: 4 x 64-bit per pixel {r, g, b, count}.  The channel sums are two's-complement fixed point
// in units of 2^-32 added with integer atomics, so the result does not depend on the order in which paths
// finish: a render is bit-reproducible, and splitting the stratum range over calls or GPUs changes nothing.
// Range +-2.1e9 per pixel and channel, resolution 2.3e-10 per sample (an fp32 radiance of 1 carries 6e-8).
// count = strata accumulated; count >= 2^48 = the pixel received a non-finite sample under RTB_FLAG_PROPAGATE_NAN
// (a bit low enough that the sum of many devices' buffers cannot carry it out of the word).
constexpr double ACCUM_SCALE = 4294967296.0;
constexpr unsigned long long ACCUM_POISON = 1ull << 48;

constexpr int BVH_STACK = 48;     // traversal stack entries (builder rejects deeper trees)
constexpr int BVH_MAX_LEAF = 4;   // primitives per leaf
// Leaf reference (negative child index): ~(first << 5 | moving << 4 | quad << 3 | count - 1).  The two kind
// bits describe the leaf's primitive when count == 1 (the builder's default), so the traversal can run
// its test without touching prim_info -- one L1 wavefront less per test on a saturated L1 data pipe.
constexpr int LEAF_KIND_QUAD = 1, LEAF_KIND_MOVING = 2;
// Both bits set (quads never move): the leaf is an axis-aligned make_box -- count = 6 quads in face order -x +x -y +y -z +z,
// the box bounds in the DPre slot of its first quad (DBoxBounds).  One slab test names the face a ray can hit first.
constexpr int LEAF_KIND_BOX = 3;
#if defined(__CUDACC__)
#define RTB_HD __host__ __device__ inline
#else
#define RTB_HD inline
#endif
RTB_HD int leaf_make(int first, int count, int kind_bits) { return ~((first << 5) | (kind_bits << 3) | (count - 1)); }
RTB_HD int leaf_first(int ref) { return (~ref) >> 5; }
RTB_HD int leaf_count(int ref) { return ((~ref) & 7) + 1; }
RTB_HD int leaf_kind_bits(int ref) { return ((~ref) >> 3) & 3; }

// primitive payload, 16 doubles (8 x double2 = 128 B), the fields of the reference structs:
//   SPHERE: cx cy | cz r | cvx cvy | cvz - | ...                 (center(t) = c + t * cv, src/object.rs:74-80)
//   QUAD  : nx ny | nz d | qx qy | qz ux | uy uz | vx vy | vz wx | wy wz     (src/object.rs:415-425:
//           n = unit normal, d = n.q, w = n/(n.n)); the test evaluates Quad::hit operation by operation
constexpr int PRIM_D2 = 8;
constexpr int PRIM_DOUBLES = 2 * PRIM_D2;

// prefilter record of a QUAD (64 B, two sectors; zero for spheres): what the conservative fp32 test of the
// wavefront traversal reads instead of the 96 cold payload bytes u | v | w (rtb_device.cuh, prefilter_quad)
struct alignas(32) DPre {
  double qx, qy, qz;   // corner q (f64: o - q cancels)
  float ax, ay, az;    // A = v x w:  alpha = w . (h x v) = A . h
  float bx, by, bz;    // B = w x u:  beta  = w . (u x h) = B . h
  float nx, ny, nz;    // unit normal, fp32
  float ab1;           // max(|A|_1, |B|_1), rounded up: scales the error bound of alpha / beta
};
static_assert(sizeof(DPre) == 64, "DPre is read as one 256-bit f64 load + one 256-bit f32 load");
struct alignas(32) DBoxBounds { double lo[3], hi[3], pad[2]; };  // in the DPre slot of a box leaf's first quad
static_assert(sizeof(DBoxBounds) == sizeof(DPre), "DBoxBounds aliases a DPre slot");

struct DMaterial {
  int kind;
  int texture;
  int needs_uv;  // texture tree contains an IMAGE node: hit u,v must be computed
  int pad;
  float color[3];
  float param;
};

struct DTexture {
  int kind;
  int a, b;       // CHECKER: even/odd texture; IMAGE: texel byte offset; NOISE: perlin table index
  int width, height;
  float color[3];
  double scale;   // CHECKER: inv_scale; NOISE: scale
};

struct DMedium {
  int first_prim, n_prims;  // boundary primitives (in `prims`, after the surfaces), DFS order
  int material;
  int cls_fast;  // bits 0..3: shading class of the phase function; bit 8: the boundary is one static sphere; bit 9: quads only; bit 10: an oriented box (obb_*)
  double neg_inv_density;
  float lo[3], hi[3];       // padded fp32 box of the boundary (line cull)
  float diag;               // diagonal of that box: no chord of the boundary is longer
  float pad;
  float sphere[4];          // boundary = one static sphere (cls_fast bit 8): centre, and r^2 shrunk by 1e-3 relative (inside test)
  // boundary = the six quads of a make_box under any Translate / RotateY chain (cls_fast bit 10): its oriented box.
  // fp32 and conservative: half_out is inflated, half_in deflated by 1e-5 x scene magnitude (device: medium_obb)
  float obb_c[3], obb_ax[3][3], obb_half_out[3], obb_half_in[3];
  float pad2[2];
};

struct alignas(32) DLight {
  double prim[16];  // same payload as `prims` (world space, time 0); first + 32-aligned: read as double2 pairs
  double q[3], u[3], v[3];  // QUAD: sampling frame
  double area;
  int kind;
  int pad;
};
static_assert(sizeof(DLight) % 32 == 0, "DLight records must keep prim[] 32-byte aligned (256-bit loads)");

// Sun::new (reference src/object.rs:223-231): unit direction, albedo, limit = 1 - angular_diameter / 180
struct DSun { float dir[3], albedo[3], limit, pad; };
constexpr int MAX_SUNS = 4;

struct DCamera {
  double center[3], pixel00[3], du[3], dv[3], disk_u[3], disk_v[3];
  double recip_sqrt_spp;
  int width, height, sqrt_spp, spp, max_depth, defocus;
  float background[3];
  float pad;
};

struct DScene {
  const float4* nodes;
  const uint4* qnodes;        // 2 per inner node: the same tree with 16-bit boxes (opt-in arm of wavefront extend)
  const float4* nodes4;       // 8 per BVH4 node: every other level of the tree collapsed (wavefront extend)
  const double2* prims;
  const int4* prim_info;
  const DPre* pre;            // one per primitive (BVH order), see DPre
  const double2* xforms;
  const DMedium* media;
  const DMaterial* materials;
  const DTexture* textures;
  const uint8_t* texels;
  const float4* perlin_vec;   // 256 per table
  const uint8_t* perlin_perm; // 768 per table: perm_x | perm_y | perm_z
  const DLight* lights;
  int n_nodes, n_surface_prims, n_prims, n_media, n_lights;
  int n_materials, n_textures;
  int bvh_depth;
  uint32_t flags;
  uint32_t seed_lo, seed_hi;
  double grid_base[3], grid_inv_cell[3];  // quantisation grid of qnodes: cell index = (x - base) * inv_cell
  float grid_cell[3];
  int use_qnodes;  // wavefront extend traverses qnodes (else nodes)
  int use_bvh4;    // wavefront extend traverses nodes4
  int defer_ok;    // every non-solid texture hangs off a Lambertian surface material (classes LAMBERT_TEX / NOISE)
  int multi_leaf;  // some BVH leaf holds more than one primitive (only with RTB_BVH_LEAF > 1)
  float scene_mag; // largest |coordinate| of the scene (primitives, media boundaries, camera): scales the f64 rounding bound of the prefilter
  int spec_bits;   // SPEC_* features the scene uses: the wavefront shade kernel picks the smallest instantiation covering them
  DCamera cam;
  int n_suns;      // > 0 only with RTB_FLAG_SUN_LIGHT (the sun term is commented out at HEAD, src/render.rs:300-308)
  DSun suns[MAX_SUNS];
};

struct DStats {
  unsigned long long paths, segments, node_visits, prim_tests, medium_probes, nonfinite;
  unsigned long long exact_tests;  // f64 reference-order tests run on candidates by the shade stage
  unsigned long long overflows;    // rays re-traced exactly because their candidates did not fit the slots
};

}  // namespace rtb
