// oracle/oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// A CPU restatement, in f64 and in the reference's own structure (recursive ray_color, list
// scans, enum dispatch, reference-shaped BVH), of the per-pixel integration path of
// carlosconley/surely-raytracing.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; librtb200.so never does.
//
// PARITY STATUS: *unpinned at the RNG boundary*.  The reference cannot be compiled here (no
// cargo/rustc, crates not vendored -- SURVEY.md F1) and ships no tests or golden vectors
// (SURVEY.md section 4).  What this file IS pinned against (tests/test_oracle_*.py):
//   - the expected values of the reference's only known-answer hook, Sphere::_test_uvs
//     (src/object.rs:134-141);
//   - tests/golden/ref_book3_*.npy, block means of the reference's own final_images/book3.png
//     (the one image HEAD can reproduce: cornell_box, src/main.rs:417-512), by PSNR;
//   - closed-form identities (white furnace, pdf normalisation, sRGB OETF values).
// Every function below cites the reference lines it follows.
//
// Two sampler modes:
//   ORC_SAMPLER_REF   -- sequential stream, rejection loops and draw order exactly as the reference
//                        (src/utils.rs:5-15, src/vec3.rs:184-250); the stream itself is a per-pixel
//                        xoshiro256++ because the reference's ThreadRng is OS-seeded (F6).
//   ORC_SAMPLER_KEYED -- the same distributions through the fixed-slot Philox4x32-10 draws and the
//                        direct (rejection-free) maps the CUDA path uses, so that single samples
//                        can be compared path by path.  Equivalence of the two modes is tested.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../include/rtb200.h"

namespace {

const double INF = std::numeric_limits<double>::infinity();  // src/utils.rs:3
const double PI = 3.14159265358979323846;                    // std::f64::consts::PI

// ------------------------------------------------------------------------------------------
// vec3 (src/vec3.rs)
// ------------------------------------------------------------------------------------------
struct Vec3 {
  double x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
  double dim(int n) const { return n == 0 ? x : (n == 1 ? y : z); }  // src/vec3.rs:64-71
  void set(int n, double v) { (n == 0 ? x : (n == 1 ? y : z)) = v; }  // src/vec3.rs:55-62
  double length_squared() const { return x * x + y * y + z * z; }     // src/vec3.rs:73-75
  double length() const { return std::sqrt(length_squared()); }
};
typedef Vec3 Point3;
typedef Vec3 Color;

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 operator*(double t, const Vec3& a) { return Vec3(t * a.x, t * a.y, t * a.z); }
inline Vec3 operator*(const Vec3& a, double t) { return t * a; }
inline Vec3 operator/(const Vec3& a, double t) { return Vec3(a.x / t, a.y / t, a.z / t); }  // src/vec3.rs:145-151: true division

inline double dot(const Vec3& u, const Vec3& v) { return u.x * v.x + u.y * v.y + u.z * v.z; }  // :167
inline Vec3 cross(const Vec3& u, const Vec3& v) {                                              // :171
  return Vec3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
inline Vec3 unit_vector(const Vec3& v) { return v / v.length(); }  // :179
inline Vec3 reflect(const Vec3& v, const Vec3& n) { return v - 2. * dot(v, n) * n; }  // :219
inline Vec3 refract(const Vec3& uv, const Vec3& n, double etai_over_etat) {           // :223-229
  double cos_theta = std::fmin(dot(-uv, n), 1.);
  Vec3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
  Vec3 r_out_parallel = std::sqrt(std::fabs(1.0 - r_out_perp.length_squared())) * -1. * n;
  return r_out_perp + r_out_parallel;
}

struct Ray {  // src/ray.rs
  Point3 orig;
  Vec3 dir;
  double tm;
  Ray() : tm(0) {}
  Ray(const Point3& o, const Vec3& d, double t = 0.) : orig(o), dir(d), tm(t) {}
  Point3 at(double t) const { return orig + t * dir; }  // :42-44
};

struct Interval {  // src/interval.rs
  double min, max;
  bool contains(double x) const { return min <= x && x <= max; }   // :21-23 closed
  bool surrounds(double x) const { return min < x && x < max; }    // :25-27 open
  double size() const { return max - min; }                        // :39-41
  Interval expand(double delta) const {                            // :43-50
    double padding = delta / 2.;
    return Interval{min - padding, max + padding};
  }
  static Interval from_intervals(const Interval& a, const Interval& b) {  // :52-57
    return Interval{std::fmin(a.min, b.min), std::fmax(a.max, b.max)};
  }
};
const Interval EMPTY = {INF, -INF};
const Interval UNIVERSE = {-INF, INF};

struct Aabb {  // src/object.rs:286-392
  Interval x, y, z;
  static Aabb empty() { return Aabb{EMPTY, EMPTY, EMPTY}; }
  static Aabb from_boxes(const Aabb& a, const Aabb& b) {  // :306-312
    return Aabb{Interval::from_intervals(a.x, b.x), Interval::from_intervals(a.y, b.y),
                Interval::from_intervals(a.z, b.z)};
  }
  static Aabb from_points(const Point3& a, const Point3& b) {  // :314-329
    return Aabb{Interval{std::fmin(a.x, b.x), std::fmax(a.x, b.x)},
                Interval{std::fmin(a.y, b.y), std::fmax(a.y, b.y)},
                Interval{std::fmin(a.z, b.z), std::fmax(a.z, b.z)}};
  }
  const Interval& axis(int n) const { return n == 0 ? x : (n == 1 ? y : z); }
  bool hit(const Ray& r, Interval ray_t) const {  // :340-370 (Q6)
    for (int a = 0; a < 3; a++) {
      double inv_d = 1. / r.dir.dim(a);
      double orig = r.orig.dim(a);
      double t0 = (axis(a).min - orig) * inv_d;
      double t1 = (axis(a).max - orig) * inv_d;
      if (inv_d < 0.) std::swap(t0, t1);
      if (t0 > ray_t.min) ray_t.min = t0;
      if (t1 < ray_t.max) ray_t.max = t1;
      if (ray_t.max <= ray_t.min) return false;
    }
    return true;
  }
  Aabb pad() const {  // :372-391
    double delta = 0.0001;
    return Aabb{x.size() >= delta ? x : x.expand(delta), y.size() >= delta ? y : y.expand(delta),
                z.size() >= delta ? z : z.expand(delta)};
  }
  Aabb shifted(const Vec3& o) const {  // impl Add<Vec3> for Aabb :394-404
    return Aabb{Interval{x.min + o.x, x.max + o.x}, Interval{y.min + o.y, y.max + o.y},
                Interval{z.min + o.z, z.max + o.z}};
  }
};

// ------------------------------------------------------------------------------------------
// RNG
// ------------------------------------------------------------------------------------------
inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants).
inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { SAMPLER_REF = 0, SAMPLER_KEYED = 1 };
const uint32_t PRIMARY_BOUNCE = 0xFFFFFFFFu;

struct Sampler {
  int mode = SAMPLER_REF;
  // REF: xoshiro256++ sequential stream
  uint64_t s[4];
  // KEYED: Philox key = seed, counter = (pixel, sample, bounce, call)
  uint32_t key[2];
  uint32_t pixel = 0, sample = 0, bounce = 0;
  uint32_t cache_call = 0xFFFFFFFFu, cache_bounce = 0, cache[4];

  void seed_ref(uint64_t seed, uint64_t stream) {
    uint64_t z = seed ^ (stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    for (int i = 0; i < 4; i++) s[i] = splitmix64(z);
  }
  uint64_t next_u64() {
    uint64_t r = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
  }
  // random_double: rand 0.8.5 Standard f64 = 53 random bits * 2^-53 in [0,1)   src/utils.rs:5-7
  double random_double() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
  double random_range(double lo, double hi) { return lo + (hi - lo) * random_double(); }  // :9-11
  int64_t random_int(int64_t lo, int64_t hi) {  // inclusive  src/utils.rs:13-15
    uint64_t span = (uint64_t)(hi - lo) + 1;
    return lo + (int64_t)(next_u64() % span);
  }
  // KEYED slot: 24-bit uniform in [0,1) -- (word >> 8) * 2^-24, identical on the CUDA side.
  double slot(uint32_t call, int word) {
    if (cache_call != call || cache_bounce != bounce) {
      uint32_t ctr[4] = {pixel, sample, bounce, call};
      philox4x32_10(ctr, key, cache);
      cache_call = call;
      cache_bounce = bounce;
    }
    return (double)(cache[word] >> 8) * (1.0 / 16777216.0);
  }
  void set_path(uint32_t px, uint32_t smp) {
    pixel = px; sample = smp; cache_call = 0xFFFFFFFFu;
  }
  void set_bounce(uint32_t b) { bounce = b; cache_call = 0xFFFFFFFFu; }
};

// src/vec3.rs:184-191 (REF) | concentric polar map (KEYED)
Vec3 random_in_unit_disk(Sampler& S) {
  if (S.mode == SAMPLER_REF) {
    for (;;) {
      double a = S.random_range(-1., 1.);
      double b = S.random_range(-1., 1.);
      Vec3 p(a, b, 0.);
      if (p.length_squared() < 1.) return p;
    }
  }
  double r = std::sqrt(S.slot(1, 0));
  double phi = 2. * PI * S.slot(1, 1);
  return Vec3(r * std::cos(phi), r * std::sin(phi), 0.);
}
// src/vec3.rs:231-238
Vec3 random_in_unit_sphere_ref(Sampler& S) {
  for (;;) {
    double a = S.random_range(-1., 1.);
    double b = S.random_range(-1., 1.);
    double c = S.random_range(-1., 1.);
    Vec3 p(a, b, c);
    if (p.length_squared() < 1.) return p;
  }
}
// src/vec3.rs:215-217 (REF) | z = 1-2 r1, phi = 2 pi r2 (KEYED): both uniform on the sphere
Vec3 random_unit_vector(Sampler& S) {
  if (S.mode == SAMPLER_REF) return unit_vector(random_in_unit_sphere_ref(S));
  double r1 = S.slot(0, 2), r2 = S.slot(0, 3);
  double z = 1. - 2. * r1;
  double rr = std::sqrt(std::fmax(0., 1. - z * z));
  double phi = 2. * PI * r2;
  return Vec3(rr * std::cos(phi), rr * std::sin(phi), z);
}
// src/vec3.rs:240-250
Vec3 random_cosine_direction(Sampler& S) {
  double r1, r2;
  if (S.mode == SAMPLER_REF) { r1 = S.random_double(); r2 = S.random_double(); }
  else { r1 = S.slot(0, 2); r2 = S.slot(0, 3); }
  double phi = 2. * PI * r1;
  double x = std::cos(phi) * std::sqrt(r2);
  double y = std::sin(phi) * std::sqrt(r2);
  double z = std::sqrt(1. - r2);
  return Vec3(x, y, z);
}

// ------------------------------------------------------------------------------------------
// ONB (src/onb.rs)
// ------------------------------------------------------------------------------------------
struct Onb {
  Vec3 axis[3];
  void build_from_w(const Vec3& w) {  // :32-47
    Vec3 unit_w = unit_vector(w);
    Vec3 a = std::fabs(unit_w.x) > 0.9 ? Vec3(0., 1., 0.) : Vec3(1., 0., 0.);
    Vec3 v = unit_vector(cross(unit_w, a));
    Vec3 u = cross(unit_w, v);
    axis[0] = u; axis[1] = v; axis[2] = unit_w;
  }
  Vec3 local(const Vec3& a) const { return a.x * axis[0] + a.y * axis[1] + a.z * axis[2]; }  // :24-30
  const Vec3& w() const { return axis[2]; }
};

// ------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------
struct Obj {
  int kind, material, first, count;
  // sphere
  Point3 center; double radius; Vec3 center_vec; bool moving;
  // quad
  Point3 q; Vec3 u, v, normal, w; double d, area;
  // transforms
  Vec3 offset; double sin_theta, cos_theta;
  // medium
  double neg_inv_density; int medium_index;
  Aabb bbox;
  int prim_id;   // canonical id (spheres/quads), -1 otherwise
  int bvh_root;  // RTB_OBJ_BVH: index into bvh_nodes, -1 if not built
};

struct BvhNode {  // src/hittable.rs:135-139
  int left, right;  // >=0: object index; <0: ~node index
  Aabb bbox;
};

struct Camera {  // src/render.rs:15-36
  int image_width, image_height, samples_per_pixel, max_depth, sqrt_spp;
  double recip_sqrt_spp, defocus_angle;
  Point3 center, pixel00_loc;
  Vec3 pixel_delta_u, pixel_delta_v, defocus_disk_u, defocus_disk_v;
  Color background;
};

struct Counters {
  uint64_t paths = 0, segments = 0, node_visits = 0, prim_tests = 0, medium_probes = 0, nonfinite = 0;
};

struct Scene {
  std::vector<Obj> objs;
  std::vector<int> children;
  std::vector<BvhNode> bvh_nodes;
  int world = -1;
  std::vector<int> lights;
  std::vector<RtbMaterial> mats;
  std::vector<RtbTexture> texs;
  struct Img { int w, h; std::vector<uint8_t> rgb; };
  std::vector<Img> images;
  std::vector<RtbPerlin> perlins;
  Camera cam;
  uint32_t flags = 0;
  uint64_t seed = 0;
  struct SunRec { Vec3 direction; Color albedo; double limit; };  // Sun  src/object.rs:216-241
  std::vector<SunRec> suns;
  int n_prims = 0, n_media = 0;
  bool use_bvh = true;
  std::string error;
};

struct HitRecord {  // src/hittable.rs:11-19
  Point3 p;
  Vec3 normal;
  int mat = -1;
  double t = 0, u = 0, v = 0;
  bool front_face = false;
  int prim = -1;  // canonical id; media: -2 - medium_index
};

inline void set_face_normal(HitRecord& rec, const Ray& r, const Vec3& outward_normal) {  // :22-37
  rec.front_face = dot(r.dir, outward_normal) < 0.;
  rec.normal = rec.front_face ? outward_normal : -outward_normal;
}

struct Ctx {
  const Scene* sc;
  Sampler S;
  Counters cnt;
};

// Camera::new  src/render.rs:62-134
int nearest_square(int i) {  // :38-41
  int r = (int)std::sqrt((double)i);
  return r * r;
}
void camera_new(const RtbCamera& c, Camera& out) {
  int image_height = (int)((double)c.image_width / c.aspect_ratio);
  if (image_height < 1) image_height = 1;
  Point3 lookfrom(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]);
  Point3 lookat(c.lookat[0], c.lookat[1], c.lookat[2]);
  Vec3 vup(c.vup[0], c.vup[1], c.vup[2]);
  Point3 center = lookfrom;
  double theta = c.vfov * (PI / 180.);  // to_radians
  double h = std::tan(theta / 2.);
  double focus_dist = c.focus_dist <= 0. ? 1. : c.focus_dist;  // :85 (Q2)
  double viewport_height = 2. * h * focus_dist;
  double viewport_width = viewport_height * (double)c.image_width / (double)image_height;
  Vec3 w = unit_vector(lookfrom - lookat);
  Vec3 u = unit_vector(cross(vup, w));
  Vec3 v = cross(w, u);
  Vec3 viewport_u = viewport_width * u;
  Vec3 viewport_v = viewport_height * -v;
  Vec3 pixel_delta_u = viewport_u / (double)c.image_width;
  Vec3 pixel_delta_v = viewport_v / (double)image_height;
  Point3 viewport_upper_left = center - (focus_dist * w) - viewport_u / 2. - viewport_v / 2.;
  Point3 pixel00_loc = viewport_upper_left + 0.5 * (pixel_delta_u + pixel_delta_v);
  double defocus_radius = focus_dist * std::tan((c.defocus_angle / 2.) * (PI / 180.));
  int spp = nearest_square(c.samples_per_pixel);  // :108 (Q1)
  double sqrt_spp = std::sqrt((double)spp);
  out.image_width = c.image_width;
  out.image_height = image_height;
  out.samples_per_pixel = spp;
  out.max_depth = c.max_depth;
  out.sqrt_spp = (int)sqrt_spp;
  out.recip_sqrt_spp = 1. / sqrt_spp;
  out.defocus_angle = c.defocus_angle;
  out.center = center;
  out.pixel00_loc = pixel00_loc;
  out.pixel_delta_u = pixel_delta_u;
  out.pixel_delta_v = pixel_delta_v;
  out.defocus_disk_u = u * defocus_radius;
  out.defocus_disk_v = v * defocus_radius;
  out.background = Color(c.background[0], c.background[1], c.background[2]);
}

// ---- bounding boxes & construction-time derived fields -------------------------------------
bool build_object(Scene& sc, int oi, std::vector<char>& seen, int depth) {
  if (oi < 0 || oi >= (int)sc.objs.size()) { sc.error = "object index out of range"; return false; }
  if (depth > 64) { sc.error = "object graph too deep"; return false; }
  if (seen[oi]) { sc.error = "object referenced twice (graph must be a tree)"; return false; }
  seen[oi] = 1;
  Obj& o = sc.objs[oi];
  o.prim_id = -1;
  o.bvh_root = -1;
  switch (o.kind) {
    case RTB_OBJ_SPHERE: {  // Sphere::new / new_moving  src/object.rs:83-105
      Vec3 rvec(o.radius, o.radius, o.radius);
      Aabb box1 = Aabb::from_points(o.center - rvec, o.center + rvec);
      if (o.moving) {
        Point3 c2 = o.center + o.center_vec;
        Aabb box2 = Aabb::from_points(c2 - rvec, c2 + rvec);
        o.bbox = Aabb::from_boxes(box1, box2);
      } else {
        o.bbox = box1;
      }
      o.prim_id = sc.n_prims++;
      break;
    }
    case RTB_OBJ_QUAD: {  // Quad::new  src/object.rs:428-445
      o.bbox = Aabb::from_points(o.q, o.q + o.u + o.v).pad();
      Vec3 n = cross(o.u, o.v);
      o.normal = unit_vector(n);
      o.w = n / dot(n, n);
      o.d = dot(o.normal, o.q);
      o.area = n.length();
      o.prim_id = sc.n_prims++;
      break;
    }
    case RTB_OBJ_LIST:
    case RTB_OBJ_BVH: {  // HittableList::add  src/hittable.rs:74-80
      if (o.first < 0 || o.count < 0 || o.first + o.count > (int)sc.children.size()) {
        sc.error = "list child range out of bounds";
        return false;
      }
      Aabb box = Aabb::empty();
      for (int k = 0; k < o.count; k++) {
        int ci = sc.children[o.first + k];
        if (!build_object(sc, ci, seen, depth + 1)) return false;
        box = Aabb::from_boxes(box, sc.objs[ci].bbox);
      }
      sc.objs[oi].bbox = box;
      break;
    }
    case RTB_OBJ_TRANSLATE: {  // Translate::new  src/transform.rs:43-53
      if (!build_object(sc, o.first, seen, depth + 1)) return false;
      sc.objs[oi].bbox = sc.objs[sc.objs[oi].first].bbox.shifted(sc.objs[oi].offset);
      break;
    }
    case RTB_OBJ_ROTATE_Y: {  // RotateY::new  src/transform.rs:143-186
      if (!build_object(sc, o.first, seen, depth + 1)) return false;
      Obj& me = sc.objs[oi];
      const Aabb& bbox = sc.objs[me.first].bbox;
      Point3 mn(INF, INF, INF), mx(-INF, -INF, -INF);
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
          for (int k = 0; k < 2; k++) {
            double fi = i, fj = j, fk = k;
            double x = fi * bbox.x.max + (1. - fi) * bbox.x.min;
            double y = fj * bbox.y.max + (1. - fj) * bbox.y.min;
            double z = fk * bbox.z.max + (1. - fk) * bbox.z.min;
            double newx = me.cos_theta * x + me.sin_theta * z;
            double newz = -me.sin_theta * x + me.cos_theta * z;
            Vec3 tester(newx, y, newz);
            for (int c = 0; c < 3; c++) {
              mn.set(c, std::fmin(mn.dim(c), tester.dim(c)));
              mx.set(c, std::fmax(mx.dim(c), tester.dim(c)));
            }
          }
      me.bbox = Aabb::from_points(mn, mx);
      break;
    }
    case RTB_OBJ_MEDIUM: {  // ConstantMedium::new  src/constant_medium.rs:23-29
      sc.objs[oi].medium_index = sc.n_media++;
      if (!build_object(sc, o.first, seen, depth + 1)) return false;
      sc.objs[oi].bbox = sc.objs[sc.objs[oi].first].bbox;  // :97-99
      break;
    }
    default:
      sc.error = "unknown object kind";
      return false;
  }
  return true;
}

// BvhNode::new  src/hittable.rs:147-187 (random axis, stable sort by bbox min, median split,
// span 1 duplicates the object, span 2 orders the pair)
int bvh_new(Scene& sc, std::vector<int>& objects, size_t start, size_t end, Sampler& rng) {
  int axis = (int)rng.random_int(0, 2);
  auto less = [&](int a, int b) {
    return sc.objs[a].bbox.axis(axis).min < sc.objs[b].bbox.axis(axis).min;
  };
  size_t span = end - start;
  int left, right;
  if (span == 1) {
    left = right = objects[start];
  } else if (span == 2) {
    if (less(objects[start], objects[start + 1])) { left = objects[start]; right = objects[start + 1]; }
    else { left = objects[start + 1]; right = objects[start]; }
  } else {
    std::stable_sort(objects.begin() + start, objects.begin() + end, less);
    size_t mid = start + span / 2;
    int l = bvh_new(sc, objects, start, mid, rng);
    int r = bvh_new(sc, objects, mid, end, rng);
    left = ~l;
    right = ~r;
  }
  auto box_of = [&](int ref) -> const Aabb& { return ref >= 0 ? sc.objs[ref].bbox : sc.bvh_nodes[~ref].bbox; };
  BvhNode n;
  n.left = left;
  n.right = right;
  n.bbox = Aabb::from_boxes(box_of(left), box_of(right));
  sc.bvh_nodes.push_back(n);
  return (int)sc.bvh_nodes.size() - 1;
}

// ---- textures (src/texture.rs, src/perlin.rs, src/rt_image.rs) ------------------------------
inline int rust_f64_as_i32(double x) {  // `as i32`: saturating, NaN -> 0
  if (!(x == x)) return 0;
  if (x >= 2147483647.0) return 2147483647;
  if (x <= -2147483648.0) return (-2147483647 - 1);
  return (int)x;
}
inline uint32_t rust_f64_as_u32(double x) {  // `as u32`: saturating, NaN -> 0
  if (!(x == x) || x <= 0.) return 0;
  if (x >= 4294967295.0) return 4294967295u;
  return (uint32_t)x;
}

double perlin_trilinear_interp(const Vec3 c[2][2][2], double u, double w, double v) {  // perlin.rs:74-96
  double uu = u * u * (3. - 2. * u);
  double vv = v * v * (3. - 2. * v);
  double ww = w * w * (3. - 2. * w);
  double accum = 0.;
  for (int _i = 0; _i < 2; _i++)
    for (int _j = 0; _j < 2; _j++)
      for (int _k = 0; _k < 2; _k++) {
        double i = _i, j = _j, k = _k;
        Vec3 weight_v(u - i, v - j, w - k);
        accum += (i * uu + (1. - i) * (1. - uu)) * (j * vv + (1. - j) * (1. - vv)) *
                 (k * ww + (1. - k) * (1. - ww)) * dot(c[_i][_j][_k], weight_v);
      }
  return accum;
}
double perlin_noise(const RtbPerlin& P, const Point3& p) {  // perlin.rs:30-54
  double u = p.x - std::floor(p.x);
  double v = p.y - std::floor(p.y);
  double w = p.z - std::floor(p.z);
  int i = rust_f64_as_i32(std::floor(p.x));
  int j = rust_f64_as_i32(std::floor(p.y));
  int k = rust_f64_as_i32(std::floor(p.z));
  Vec3 c[2][2][2];
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++)
      for (int dk = 0; dk < 2; dk++) {
        int idx = P.perm_x[(i + di) & 255] ^ P.perm_y[(j + dj) & 255] ^ P.perm_z[(k + dk) & 255];
        c[di][dj][dk] = Vec3(P.ranvec[idx][0], P.ranvec[idx][1], P.ranvec[idx][2]);
      }
  return perlin_trilinear_interp(c, u, w, v);
}
double perlin_turb(const RtbPerlin& P, const Point3& p, int depth = 7) {  // perlin.rs:56-72
  double accum = 0.;
  Point3 temp_p = p;
  double weight = 1.;
  for (int i = 0; i < depth; i++) {
    accum += weight * perlin_noise(P, temp_p);
    weight *= 0.5;
    temp_p = temp_p * 2.;
  }
  return std::fabs(accum);
}

Color texture_value(const Scene& sc, int ti, double u, double v, const Point3& p, int depth = 0) {
  if (ti < 0 || ti >= (int)sc.texs.size() || depth > 16) return Color(0., 0., 0.);
  const RtbTexture& t = sc.texs[ti];
  switch (t.kind) {
    case RTB_TEX_SOLID:  // texture.rs:44-46
      return Color(t.color[0], t.color[1], t.color[2]);
    case RTB_TEX_CHECKER: {  // texture.rs:71-81 (Q19: Rust % keeps the sign; x+y+z wraps in release)
      int x = rust_f64_as_i32(std::floor(t.scale * p.x));
      int y = rust_f64_as_i32(std::floor(t.scale * p.y));
      int z = rust_f64_as_i32(std::floor(t.scale * p.z));
      int32_t s = (int32_t)((uint32_t)x + (uint32_t)y + (uint32_t)z);
      return texture_value(sc, (s % 2 == 0) ? t.a : t.b, u, v, p, depth + 1);
    }
    case RTB_TEX_IMAGE: {  // texture.rs:95-107 + rt_image.rs:37-46 (Q20)
      const Scene::Img& im = sc.images[t.a];
      if (im.h <= 0) return Color(0., 1., 1.);
      double uc = std::fmin(std::fmax(u, 0.), 1.);
      double vc = std::fmin(std::fmax(v, 0.), 1.);
      uint32_t i = rust_f64_as_u32(uc * (double)im.w);
      uint32_t j = rust_f64_as_u32(vc * (double)im.h);
      uint32_t x = std::min(i, (uint32_t)im.w - 1);
      uint32_t y = (uint32_t)im.h - j - 1;  // wraps when j == h (release build), then clamps
      y = std::min(y, (uint32_t)im.h - 1);
      const uint8_t* px = &im.rgb[3 * ((size_t)y * im.w + x)];
      double color_scale = 1.0 / 255.0;
      return Color(px[0] * color_scale, px[1] * color_scale, px[2] * color_scale);
    }
    case RTB_TEX_NOISE: {  // texture.rs:127-130 (Q21)
      Vec3 s = t.scale * p;
      return Color(1., 1., 1.) * 0.5 * (1. + std::sin(s.z + 10. * perlin_turb(sc.perlins[t.a], s)));
    }
  }
  return Color(0., 0., 0.);
}

// ---- geometry ------------------------------------------------------------------------------
bool hit_object(Ctx& C, int oi, const Ray& r, const Interval& ray_t, HitRecord& rec);

// get_sphere_uv  src/object.rs:114-120 (Q8)
void get_sphere_uv(const Point3& p, double& u, double& v) {
  double theta = std::acos(-p.y);
  double phi = std::atan2(-p.z, p.x) + PI;
  double inv_pi = 1.0 / PI;
  u = phi * inv_pi * 0.5;
  v = theta * inv_pi;
}

bool sphere_hit(Ctx& C, const Obj& s, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // :145-184
  C.cnt.prim_tests++;
  Point3 center = s.moving ? s.center + r.tm * s.center_vec : s.center;  // :107-112
  Vec3 oc = r.orig - center;
  double a = r.dir.length_squared();
  double half_b = dot(oc, r.dir);
  double c = oc.length_squared() - s.radius * s.radius;
  double discriminant = half_b * half_b - a * c;
  if (discriminant < 0.) return false;
  double sqrtd = std::sqrt(discriminant);
  double root = (-half_b - sqrtd) / a;
  if (!ray_t.surrounds(root)) {
    root = (sqrtd - half_b) / a;
    if (!ray_t.surrounds(root)) return false;
  }
  rec.t = root;
  rec.p = r.at(root);
  Vec3 outward_normal = (rec.p - center) / s.radius;
  get_sphere_uv(outward_normal, rec.u, rec.v);
  rec.mat = s.material;
  rec.prim = s.prim_id;
  set_face_normal(rec, r, outward_normal);
  return true;
}

bool quad_hit(Ctx& C, const Obj& q, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // :453-490
  C.cnt.prim_tests++;
  double denom = dot(q.normal, r.dir);
  if (std::fabs(denom) < 1e-8) return false;
  double t = (q.d - dot(q.normal, r.orig)) / denom;
  if (!ray_t.contains(t)) return false;
  Point3 intersection = r.at(t);
  Vec3 planar_hitpt_vector = intersection - q.q;
  double a = dot(q.w, cross(planar_hitpt_vector, q.v));
  double b = dot(q.w, cross(q.u, planar_hitpt_vector));
  if ((a < 0.) || (1. < a) || (b < 0.) || (1. < b)) return false;
  rec.t = t;
  rec.p = intersection;
  rec.mat = q.material;
  rec.u = a;
  rec.v = b;
  rec.prim = q.prim_id;
  set_face_normal(rec, r, q.normal);
  return true;
}

bool list_hit(Ctx& C, const Obj& l, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // hittable.rs:88-109
  bool any = false;
  double closest_so_far = ray_t.max;
  HitRecord temp;
  for (int k = 0; k < l.count; k++) {
    if (hit_object(C, C.sc->children[l.first + k], r, Interval{ray_t.min, closest_so_far}, temp)) {
      closest_so_far = temp.t;
      rec = temp;
      any = true;
    }
  }
  return any;
}

bool bvh_ref_hit(Ctx& C, int ref, const Ray& r, const Interval& ray_t, HitRecord& rec);
bool bvh_node_hit(Ctx& C, const BvhNode& n, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // :216-236
  C.cnt.node_visits++;
  if (!n.bbox.hit(r, ray_t)) return false;
  HitRecord left;
  if (bvh_ref_hit(C, n.left, r, ray_t, left)) {
    HitRecord right;
    if (bvh_ref_hit(C, n.right, r, Interval{ray_t.min, left.t}, right)) rec = right;
    else rec = left;
    return true;
  }
  return bvh_ref_hit(C, n.right, r, ray_t, rec);
}
bool bvh_ref_hit(Ctx& C, int ref, const Ray& r, const Interval& ray_t, HitRecord& rec) {
  if (ref >= 0) return hit_object(C, ref, r, ray_t, rec);
  return bvh_node_hit(C, C.sc->bvh_nodes[~ref], r, ray_t, rec);
}

bool translate_hit(Ctx& C, const Obj& t, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // transform.rs:57-69
  Ray offset_r(r.orig - t.offset, r.dir, r.tm);
  if (!hit_object(C, t.first, offset_r, ray_t, rec)) return false;
  rec.p = rec.p + t.offset;
  return true;
}

bool rotate_y_hit(Ctx& C, const Obj& o, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // :85-135 (Q18)
  Point3 origin = r.orig;
  Vec3 direction = r.dir;
  origin.x = o.cos_theta * r.orig.x - o.sin_theta * r.orig.z;
  origin.z = o.sin_theta * r.orig.x + o.cos_theta * r.orig.z;
  direction.x = o.cos_theta * r.dir.x - o.sin_theta * r.dir.z;
  direction.z = o.sin_theta * r.dir.x + o.cos_theta * r.dir.z;
  Ray rotated_r(origin, direction, r.tm);
  if (!hit_object(C, o.first, rotated_r, ray_t, rec)) return false;
  Point3 p = rec.p;
  p.x = o.cos_theta * rec.p.x + o.sin_theta * rec.p.z;
  p.z = -o.sin_theta * rec.p.x + o.cos_theta * rec.p.z;
  Vec3 normal = rec.normal;
  normal.x = o.cos_theta * rec.normal.x + o.sin_theta * rec.normal.z;
  normal.z = -o.sin_theta * rec.normal.x + o.cos_theta * rec.normal.z;
  rec.p = p;
  rec.normal = normal;
  return true;
}

// boundary probes of ConstantMedium::hit  src/constant_medium.rs:46-55
bool medium_interval(Ctx& C, const Obj& m, const Ray& r, double& t1, double& t2) {
  C.cnt.medium_probes++;
  HitRecord rec1, rec2;
  if (!hit_object(C, m.first, r, UNIVERSE, rec1)) return false;
  if (!hit_object(C, m.first, r, Interval{rec1.t + 0.0001, INF}, rec2)) return false;
  t1 = rec1.t;
  t2 = rec2.t;
  return true;
}

bool medium_hit(Ctx& C, const Obj& m, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // :41-95 (Q17)
  double t1, t2;
  if (!medium_interval(C, m, r, t1, t2)) return false;
  if (t1 < ray_t.min) t1 = ray_t.min;
  if (t2 > ray_t.max) t2 = ray_t.max;
  if (t1 >= t2) return false;
  if (t1 < 0.) t1 = 0.;
  double ray_length = r.dir.length();
  double distance_inside_boundary = (t2 - t1) * ray_length;
  double U = (C.S.mode == SAMPLER_REF)
                 ? C.S.random_double()
                 : C.S.slot(1u + (uint32_t)m.medium_index / 4u, m.medium_index % 4);
  double hit_distance = m.neg_inv_density * std::log(U);
  if (hit_distance > distance_inside_boundary) return false;
  double t = t1 + hit_distance / ray_length;
  rec.t = t;
  rec.p = r.at(t);
  rec.normal = Vec3(1., 0., 0.);
  rec.front_face = true;
  rec.mat = m.material;
  rec.u = 0.;
  rec.v = 0.;
  rec.prim = -2 - m.medium_index;
  return true;
}

bool hit_object(Ctx& C, int oi, const Ray& r, const Interval& ray_t, HitRecord& rec) {  // object.rs:29-39
  const Obj& o = C.sc->objs[oi];
  switch (o.kind) {
    case RTB_OBJ_SPHERE: return sphere_hit(C, o, r, ray_t, rec);
    case RTB_OBJ_QUAD: return quad_hit(C, o, r, ray_t, rec);
    case RTB_OBJ_LIST: return list_hit(C, o, r, ray_t, rec);
    case RTB_OBJ_BVH:
      if (o.bvh_root >= 0 && C.sc->use_bvh) return bvh_node_hit(C, C.sc->bvh_nodes[o.bvh_root], r, ray_t, rec);
      return list_hit(C, o, r, ray_t, rec);
    case RTB_OBJ_TRANSLATE: return translate_hit(C, o, r, ray_t, rec);
    case RTB_OBJ_ROTATE_Y: return rotate_y_hit(C, o, r, ray_t, rec);
    case RTB_OBJ_MEDIUM: return medium_hit(C, o, r, ray_t, rec);
  }
  return false;
}

// ---- light sampling (src/object.rs:122-132,190-212,492-506; src/hittable.rs:115-129) --------
double object_pdf_value(Ctx& C, int oi, const Point3& origin, const Vec3& direction) {  // object.rs:62-69
  const Obj& o = C.sc->objs[oi];
  switch (o.kind) {
    case RTB_OBJ_QUAD: {  // :492-501 (Q11)
      HitRecord rec;
      if (!quad_hit(C, o, Ray(origin, direction), Interval{0.001, INF}, rec)) return 0.;
      double distance_squared = rec.t * rec.t * direction.length_squared();
      double cosine = std::fabs(dot(direction, rec.normal) / direction.length());
      return distance_squared / (cosine * o.area);
    }
    case RTB_OBJ_SPHERE: {  // :190-202 (Q9, Q10)
      HitRecord rec;
      if (!sphere_hit(C, o, Ray(origin, direction), Interval{0.001, INF}, rec)) return 0.;
      double cos_theta_max = std::sqrt(1. - o.radius * o.radius / (o.center - origin).length_squared());
      double solid_angle = 2. * PI * (1. - cos_theta_max);
      return 1. / solid_angle;
    }
    case RTB_OBJ_LIST: {  // hittable.rs:115-124 (Q12); an empty list panics in the reference
      if (o.count == 0) return std::numeric_limits<double>::quiet_NaN();
      double weight = 1. / (double)o.count;
      double sum = 0.;
      for (int k = 0; k < o.count; k++) sum += object_pdf_value(C, C.sc->children[o.first + k], origin, direction);
      return sum * weight;
    }
    default:
      return 0.;
  }
}

Vec3 object_random(Ctx& C, int oi, const Point3& origin) {  // object.rs:53-60
  const Obj& o = C.sc->objs[oi];
  Sampler& S = C.S;
  switch (o.kind) {
    case RTB_OBJ_QUAD: {  // :503-506
      double r1, r2;
      if (S.mode == SAMPLER_REF) { r1 = S.random_double(); r2 = S.random_double(); }
      else { r1 = S.slot(0, 2); r2 = S.slot(0, 3); }
      Point3 p = o.q + (r1 * o.u) + (r2 * o.v);
      return p - origin;
    }
    case RTB_OBJ_SPHERE: {  // :204-212 + random_to_sphere :122-132
      Vec3 direction = o.center - origin;
      double distance_squared = direction.length_squared();
      Onb uvw;
      uvw.build_from_w(direction);
      double r1, r2;
      if (S.mode == SAMPLER_REF) { r1 = S.random_double(); r2 = S.random_double(); }
      else { r1 = S.slot(0, 2); r2 = S.slot(0, 3); }
      double z = 1. + r2 * (std::sqrt(1. - o.radius * o.radius / distance_squared) - 1.);
      double phi = 2. * PI * r1;
      double x = std::cos(phi) * std::sqrt(1. - z * z);
      double y = std::sin(phi) * std::sqrt(1. - z * z);
      return uvw.local(Vec3(x, y, z));
    }
    case RTB_OBJ_LIST: {  // hittable.rs:126-129 (nested lists in KEYED mode reuse the pick slot)
      if (o.count == 0) return Vec3(1., 0., 0.);
      int64_t pick;
      if (S.mode == SAMPLER_REF) pick = S.random_int(0, (int64_t)o.count - 1);
      else pick = std::min<int64_t>((int64_t)o.count - 1, (int64_t)(S.slot(0, 1) * (double)o.count));
      return object_random(C, C.sc->children[o.first + (int)pick], origin);
    }
    default:
      return Vec3(1., 0., 0.);
  }
}

// the `lights` argument: Object::List(lights)  (src/main.rs:485-494, src/render.rs:141)
double lights_pdf_value(Ctx& C, const Point3& origin, const Vec3& direction) {
  const std::vector<int>& L = C.sc->lights;
  double weight = 1. / (double)L.size();
  double sum = 0.;
  for (size_t k = 0; k < L.size(); k++) sum += object_pdf_value(C, L[k], origin, direction);
  return sum * weight;
}
Vec3 lights_random(Ctx& C, const Point3& origin) {
  const std::vector<int>& L = C.sc->lights;
  int64_t n = (int64_t)L.size();
  int64_t pick;
  if (C.S.mode == SAMPLER_REF) pick = C.S.random_int(0, n - 1);
  else pick = std::min<int64_t>(n - 1, (int64_t)(C.S.slot(0, 1) * (double)n));
  return object_random(C, L[(size_t)pick], origin);
}

// ---- materials (src/material.rs) -------------------------------------------------------------
enum { SREC_NONE = 0, SREC_SKIP = 1, SREC_PDF_COSINE = 2, SREC_PDF_SPHERE = 3 };
struct ScatterRecord {  // :15-23
  Color attenuation;
  int kind = SREC_NONE;
  Ray skip_ray;
  Onb uvw;  // CosinePDF
};

double dielectric_reflectance(double cosine, double ref_idx) {  // :156-163
  double r0 = (1. - ref_idx) / (1. + ref_idx);
  r0 = r0 * r0;
  return r0 + (1. - r0) * std::pow(1. - cosine, 5.);
}

void material_scatter(Ctx& C, const RtbMaterial& m, const Ray& r_in, const HitRecord& rec, ScatterRecord& srec) {
  switch (m.kind) {
    case RTB_MAT_LAMBERTIAN:  // :93-98
      srec.attenuation = texture_value(*C.sc, m.texture, rec.u, rec.v, rec.p);
      srec.kind = SREC_PDF_COSINE;
      srec.uvw.build_from_w(rec.normal);  // CosinePDF::new  pdf.rs:60-66
      return;
    case RTB_MAT_METAL: {  // :125-134 (Q14)
      Vec3 reflected = reflect(unit_vector(r_in.dir), rec.normal);
      reflected = unit_vector(reflected) + (m.param * random_unit_vector(C.S));
      srec.attenuation = Color(m.color[0], m.color[1], m.color[2]);
      srec.kind = SREC_SKIP;
      srec.skip_ray = Ray(rec.p, reflected, r_in.tm);
      return;
    }
    case RTB_MAT_DIELECTRIC: {  // :167-191 (Q15)
      double refraction_ratio = rec.front_face ? 1.0 / m.param : m.param;
      Vec3 unit_direction = unit_vector(r_in.dir);
      double cos_theta = std::fmin(dot(-unit_direction, rec.normal), 1.);
      double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
      bool cannot_refract = refraction_ratio * sin_theta > 1.0;
      bool do_reflect = cannot_refract;
      if (!do_reflect) {  // short-circuit ||: the draw happens only when refraction is possible
        double U = (C.S.mode == SAMPLER_REF) ? C.S.random_double() : C.S.slot(0, 0);
        do_reflect = dielectric_reflectance(cos_theta, refraction_ratio) > U;
      }
      Vec3 direction = do_reflect ? reflect(unit_direction, rec.normal)
                                  : refract(unit_direction, rec.normal, refraction_ratio);
      srec.attenuation = Color(m.color[0], m.color[1], m.color[2]);
      srec.kind = SREC_SKIP;
      srec.skip_ray = Ray(rec.p, direction, r_in.tm);
      return;
    }
    case RTB_MAT_DIFFUSE_LIGHT:  // :218-222
      srec.kind = SREC_NONE;
      return;
    case RTB_MAT_ISOTROPIC:  // :241-248
      srec.attenuation = texture_value(*C.sc, m.texture, rec.u, rec.v, rec.p);
      srec.kind = SREC_PDF_SPHERE;
      return;
  }
  srec.kind = SREC_NONE;
}

Color material_emitted(Ctx& C, const RtbMaterial& m, const HitRecord& rec) {  // :45-50, 210-215 (Q16)
  if (m.kind == RTB_MAT_DIFFUSE_LIGHT && rec.front_face) return texture_value(*C.sc, m.texture, rec.u, rec.v, rec.p);
  return Color(0., 0., 0.);
}

double material_scattering_pdf(Ctx& C, const RtbMaterial& m, const HitRecord& rec, const Ray& scattered) {  // :52-58
  if (m.kind == RTB_MAT_LAMBERTIAN) {  // :100-109
    double cos_theta = dot(rec.normal, unit_vector(scattered.dir));
    return cos_theta < 0. ? 0. : cos_theta / PI;
  }
  if (m.kind == RTB_MAT_ISOTROPIC)  // F3: HEAD has the trait default 0; intent is 1/(4 pi)
    return (C.sc->flags & RTB_FLAG_ISO_PDF_ZERO) ? 0. : 1. / (4. * PI);
  return 0.;
}

// pdf.rs: SpherePDF :44-54, CosinePDF :56-78
double material_pdf_value(const ScatterRecord& srec, const Vec3& direction) {
  if (srec.kind == SREC_PDF_SPHERE) return 1. / (4. * PI);
  double cosine_theta = dot(unit_vector(direction), srec.uvw.w());
  return std::fmax(0., cosine_theta / PI);
}
Vec3 material_pdf_generate(Ctx& C, const ScatterRecord& srec) {
  if (srec.kind == SREC_PDF_SPHERE) return random_unit_vector(C.S);
  return srec.uvw.local(random_cosine_direction(C.S));
}

// ---- integrator (src/render.rs) --------------------------------------------------------------
inline bool finite3(const Color& c) { return std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z); }

Color ray_color(Ctx& C, const Ray& r, int depth) {  // :251-312
  const Scene& sc = *C.sc;
  if (depth <= 0) return Color(0., 0., 0.);
  C.S.set_bounce((uint32_t)(sc.cam.max_depth - depth));
  C.cnt.segments++;
  HitRecord rec;
  if (!hit_object(C, sc.world, r, Interval{0.0001, INF}, rec)) {  // :264-270, 298-309
    // HEAD: `cam.background //+ sun_light` -- the sun term is commented out (Q23).  RTB_FLAG_SUN_LIGHT restates the
    // commented lines :300-306: sun_light = sum over suns of Sun::_hit(r) (object.rs:232-239)
    Color sun_light(0., 0., 0.);
    if (sc.flags & RTB_FLAG_SUN_LIGHT) {
      const Vec3 unit_direction = unit_vector(r.dir);
      for (const Scene::SunRec& sun : sc.suns)
        if (dot(unit_direction, sun.direction) > sun.limit) sun_light = sun_light + sun.albedo;
    }
    return sc.cam.background + sun_light;
  }
  const RtbMaterial& mat = sc.mats[rec.mat];
  Color color_from_emission = material_emitted(C, mat, rec);  // :272
  ScatterRecord srec;
  material_scatter(C, mat, r, rec, srec);  // :273
  if (srec.kind == SREC_NONE) return color_from_emission;  // :295
  if (srec.kind == SREC_SKIP)  // :275-277
    return srec.attenuation * ray_color(C, srec.skip_ray, depth - 1);
  // PdfPtr :278-293.  MixturePDF(HittablePDF(lights), material pdf)  pdf.rs:102-127 (Q13);
  // F2 rule: an empty light list degenerates to the material pdf alone.
  Vec3 dir;
  double pdf_val;
  bool have_lights = !sc.lights.empty();
  if (have_lights) {
    double U = (C.S.mode == SAMPLER_REF) ? C.S.random_double() : C.S.slot(0, 0);
    dir = (U < 0.5) ? lights_random(C, rec.p) : material_pdf_generate(C, srec);
  } else {
    dir = material_pdf_generate(C, srec);
  }
  Ray scattered(rec.p, dir, r.tm);  // :283
  if (have_lights)
    pdf_val = 0.5 * lights_pdf_value(C, rec.p, scattered.dir) + 0.5 * material_pdf_value(srec, scattered.dir);
  else
    pdf_val = material_pdf_value(srec, scattered.dir);
  double scattering_pdf = material_scattering_pdf(C, mat, rec, scattered);  // :286
  if (!(sc.flags & RTB_FLAG_PROPAGATE_NAN)) {
    // default NaN policy (Q22): a zero / non-finite pdf makes the sample contribute nothing
    if (!(pdf_val > 0.) || !std::isfinite(pdf_val)) {
      C.cnt.nonfinite++;
      return color_from_emission;
    }
  }
  Color sample_color = ray_color(C, scattered, depth - 1);  // :287-288
  Color color_from_scatter = (srec.attenuation * scattering_pdf * sample_color) / pdf_val;  // :289-290
  return color_from_emission + color_from_scatter;
}

Ray get_ray(Ctx& C, int i, int j, int s_i, int s_j) {  // :218-249
  const Camera& cam = C.sc->cam;
  Sampler& S = C.S;
  Point3 pixel_center = cam.pixel00_loc + ((double)i * cam.pixel_delta_u) + ((double)j * cam.pixel_delta_v);
  double u1, u2;
  if (S.mode == SAMPLER_REF) { u1 = S.random_double(); u2 = S.random_double(); }
  else { u1 = S.slot(0, 0); u2 = S.slot(0, 1); }
  double px = -0.5 + cam.recip_sqrt_spp * ((double)s_i + u1);  // :246-247
  double py = -0.5 + cam.recip_sqrt_spp * ((double)s_j + u2);
  Point3 pixel_sample = pixel_center + (px * cam.pixel_delta_u + py * cam.pixel_delta_v);
  Point3 ray_origin = cam.center;
  if (!(cam.defocus_angle <= 0.)) {  // :226-230, 238-241
    Vec3 p = random_in_unit_disk(S);
    ray_origin = cam.center + (p.x * cam.defocus_disk_u) + (p.y * cam.defocus_disk_v);
  }
  Vec3 ray_direction = pixel_sample - ray_origin;
  double ray_time = (S.mode == SAMPLER_REF) ? S.random_double() : S.slot(0, 2);  // :233
  return Ray(ray_origin, ray_direction, ray_time);
}

// write_color  src/color.rs:8-59
double linear_to_gamma(double linear) {  // :53-59
  if (linear <= 0.0031308) return 12.92 * linear;
  return 1.055 * std::pow(linear, 1. / 2.4) - 0.055;
}
uint8_t rust_f64_as_u8(double x) {
  if (!(x == x) || x <= 0.) return 0;
  if (x >= 255.) return 255;
  return (uint8_t)x;
}

}  // namespace

// ==============================================================================================
// C interface (used through ctypes by tests/ and bench.py only)
// ==============================================================================================
extern "C" {

struct orc_scene {
  Scene sc;
};

static thread_local std::string g_err;
const char* orc_last_error(void) { return g_err.c_str(); }

int orc_scene_create(const RtbSceneDesc* d, orc_scene** out) {
  if (!d || !out) { g_err = "null argument"; return RTB_ERR_INVALID; }
  if (d->abi_version != RTB_ABI_VERSION) { g_err = "abi version mismatch"; return RTB_ERR_INVALID; }
  orc_scene* h = new orc_scene();
  Scene& sc = h->sc;
  sc.flags = d->flags;
  sc.seed = d->seed;
  for (int i = 0; i < d->n_suns; i++) {  // Sun::new  src/object.rs:223-231
    const RtbSun& u = d->suns[i];
    sc.suns.push_back(Scene::SunRec{unit_vector(Vec3(u.direction[0], u.direction[1], u.direction[2])),
                                    Color(u.albedo[0], u.albedo[1], u.albedo[2]), 1. - u.angular_diameter / 180.});
  }
  sc.objs.resize(d->n_objects);
  for (int i = 0; i < d->n_objects; i++) {
    const RtbObject& s = d->objects[i];
    Obj& o = sc.objs[i];
    o = Obj();
    o.kind = s.kind; o.material = s.material; o.first = s.first; o.count = s.count;
    o.moving = false; o.medium_index = -1; o.prim_id = -1; o.bvh_root = -1;
    switch (s.kind) {
      case RTB_OBJ_SPHERE:
        o.center = Vec3(s.v[0], s.v[1], s.v[2]); o.radius = s.v[3];
        o.center_vec = Vec3(s.v[4], s.v[5], s.v[6]); o.moving = s.v[7] != 0.;
        break;
      case RTB_OBJ_QUAD:
        o.q = Vec3(s.v[0], s.v[1], s.v[2]); o.u = Vec3(s.v[3], s.v[4], s.v[5]); o.v = Vec3(s.v[6], s.v[7], s.v[8]);
        break;
      case RTB_OBJ_TRANSLATE: o.offset = Vec3(s.v[0], s.v[1], s.v[2]); break;
      case RTB_OBJ_ROTATE_Y: {  // transform.rs:144-146
        double radians = s.v[0] * (PI / 180.);
        o.sin_theta = std::sin(radians);
        o.cos_theta = std::cos(radians);
        break;
      }
      case RTB_OBJ_MEDIUM: o.neg_inv_density = -1. / s.v[0]; break;
      default: break;
    }
  }
  sc.children.assign(d->children, d->children + d->n_children);
  sc.world = d->world;
  sc.mats.assign(d->materials, d->materials + d->n_materials);
  sc.texs.assign(d->textures, d->textures + d->n_textures);
  for (int i = 0; i < d->n_images; i++) {
    Scene::Img im;
    im.w = d->images[i].width; im.h = d->images[i].height;
    im.rgb.assign(d->images[i].rgb, d->images[i].rgb + (size_t)3 * im.w * im.h);
    sc.images.push_back(std::move(im));
  }
  sc.perlins.assign(d->perlins, d->perlins + d->n_perlins);
  camera_new(d->camera, sc.cam);

  std::vector<char> seen(sc.objs.size(), 0);
  if (sc.world < 0 || sc.world >= (int)sc.objs.size() ||
      (sc.objs[sc.world].kind != RTB_OBJ_LIST && sc.objs[sc.world].kind != RTB_OBJ_BVH)) {
    g_err = "world must be a list"; delete h; return RTB_ERR_INVALID;
  }
  if (!build_object(sc, sc.world, seen, 0)) { g_err = sc.error; delete h; return RTB_ERR_INVALID; }
  sc.lights.assign(d->lights, d->lights + d->n_lights);
  for (int li : sc.lights) {
    if (li < 0 || li >= (int)sc.objs.size()) { g_err = "light index out of range"; delete h; return RTB_ERR_INVALID; }
    if (!seen[li]) {
      int saved_prims = sc.n_prims, saved_media = sc.n_media;
      if (!build_object(sc, li, seen, 0)) { g_err = sc.error; delete h; return RTB_ERR_INVALID; }
      // light-only objects do not get canonical ids
      sc.n_prims = saved_prims; sc.n_media = saved_media;
    }
  }
  for (size_t i = 0; i < sc.objs.size(); i++) {
    const Obj& o = sc.objs[i];
    if (!seen[i]) continue;
    if ((o.kind == RTB_OBJ_SPHERE || o.kind == RTB_OBJ_QUAD || o.kind == RTB_OBJ_MEDIUM) &&
        (o.material < 0 || o.material >= (int)sc.mats.size())) {
      g_err = "material index out of range"; delete h; return RTB_ERR_INVALID;
    }
  }
  // reference-shaped BVHs where the scene called create_bvh (random axis per node, seeded)
  Sampler rng;
  rng.seed_ref(sc.seed, 0xB5ull);
  for (size_t i = 0; i < sc.objs.size(); i++) {
    if (!seen[i] || sc.objs[i].kind != RTB_OBJ_BVH || sc.objs[i].count == 0) continue;
    std::vector<int> list(sc.children.begin() + sc.objs[i].first,
                          sc.children.begin() + sc.objs[i].first + sc.objs[i].count);
    sc.objs[i].bvh_root = bvh_new(sc, list, 0, list.size(), rng);
  }
  *out = h;
  return RTB_OK;
}

void orc_scene_destroy(orc_scene* h) { delete h; }
void orc_set_use_bvh(orc_scene* h, int on) { h->sc.use_bvh = on != 0; }

int orc_scene_info(const orc_scene* h, RtbSceneInfo* info) {
  std::memset(info, 0, sizeof(*info));
  const Scene& sc = h->sc;
  info->image_width = sc.cam.image_width;
  info->image_height = sc.cam.image_height;
  info->spp_used = sc.cam.samples_per_pixel;
  info->sqrt_spp = sc.cam.sqrt_spp;
  info->max_depth = sc.cam.max_depth;
  info->n_surface_prims = sc.n_prims;
  info->n_media = sc.n_media;
  info->n_bvh_nodes = (int)sc.bvh_nodes.size();
  info->n_lights = (int)sc.lights.size();
  info->device = -1;
  return RTB_OK;
}

struct OrcStats {
  uint64_t paths, segments, node_visits, prim_tests, medium_probes, nonfinite_samples;
  double wall_ms;
  int32_t threads;
  int32_t reserved;
};

// render_par_lights  src/render.rs:144-216: 3-scanline chunks handed to worker threads (rayon's
// work stealing is replaced by an atomic chunk counter); per pixel the s_j/s_i stratum loops.
// sum[3*idx+c] += colour (Q24); sumsq (optional) accumulates squares for the variance estimate.
int orc_render(orc_scene* h, int64_t sample_begin, int64_t sample_end, int sampler_mode, int threads,
               double* sum, double* sumsq, OrcStats* stats) {
  const Scene& sc = h->sc;
  const Camera& cam = sc.cam;
  if (sample_begin < 0 || sample_end > cam.samples_per_pixel || sample_begin > sample_end) {
    g_err = "sample range out of bounds"; return RTB_ERR_INVALID;
  }
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  const int64_t n_pixels = (int64_t)cam.image_width * cam.image_height;
  const int64_t chunk_size = (int64_t)cam.image_width * 3;  // :171
  const int64_t n_chunks = (n_pixels + chunk_size - 1) / chunk_size;
  std::atomic<int64_t> next_chunk(0);
  std::vector<Counters> counters(threads);
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&](int tid) {
    Ctx C;
    C.sc = &sc;
    C.S.mode = sampler_mode;
    C.S.key[0] = (uint32_t)sc.seed;
    C.S.key[1] = (uint32_t)(sc.seed >> 32);
    for (;;) {
      int64_t j = next_chunk.fetch_add(1);
      if (j >= n_chunks) break;
      int64_t lo = j * chunk_size, hi = std::min(n_pixels, lo + chunk_size);
      for (int64_t idx = lo; idx < hi; idx++) {
        int x = (int)(idx % cam.image_width);
        int y = (int)(idx / cam.image_width);
        if (sampler_mode == SAMPLER_REF) C.S.seed_ref(sc.seed, (uint64_t)idx * 1000003ull + (uint64_t)sample_begin);
        for (int64_t s = sample_begin; s < sample_end; s++) {
          int s_j = (int)(s / cam.sqrt_spp), s_i = (int)(s % cam.sqrt_spp);  // :185-186 order
          C.S.set_path((uint32_t)idx, (uint32_t)s);
          C.S.set_bounce(PRIMARY_BOUNCE);
          Ray r = get_ray(C, x, y, s_i, s_j);
          Color color = ray_color(C, r, cam.max_depth);
          C.cnt.paths++;
          if (!(sc.flags & RTB_FLAG_PROPAGATE_NAN) && !finite3(color)) {
            C.cnt.nonfinite++;
            color = Color(0., 0., 0.);
          }
          sum[3 * idx + 0] += color.x;
          sum[3 * idx + 1] += color.y;
          sum[3 * idx + 2] += color.z;
          if (sumsq) {
            sumsq[3 * idx + 0] += color.x * color.x;
            sumsq[3 * idx + 1] += color.y * color.y;
            sumsq[3 * idx + 2] += color.z * color.z;
          }
        }
      }
    }
    counters[tid] = C.cnt;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    for (const Counters& c : counters) {
      stats->paths += c.paths; stats->segments += c.segments; stats->node_visits += c.node_visits;
      stats->prim_tests += c.prim_tests; stats->medium_probes += c.medium_probes;
      stats->nonfinite_samples += c.nonfinite;
    }
    stats->wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    stats->threads = threads;
  }
  return RTB_OK;
}

// HittableList::hit over the SURFACES only (media skipped: they are stochastic), for the
// deterministic parity harness.  Media are skipped by hiding them from the list scan.
int orc_trace(orc_scene* h, const RtbRay* rays, int64_t n, uint32_t /*flags*/, RtbHit* hits) {
  Scene sc_copy = h->sc;  // hide media: turn them into empty lists
  for (Obj& o : sc_copy.objs)
    if (o.kind == RTB_OBJ_MEDIUM) { o.kind = RTB_OBJ_LIST; o.first = 0; o.count = 0; }
  Ctx C;
  C.sc = &sc_copy;
  for (int64_t i = 0; i < n; i++) {
    Ray r(Vec3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]),
          Vec3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), rays[i].time);
    HitRecord rec;
    RtbHit& o = hits[i];
    std::memset(&o, 0, sizeof(o));
    if (hit_object(C, sc_copy.world, r, Interval{rays[i].t_min, INF}, rec)) {
      o.prim = rec.prim; o.front_face = rec.front_face ? 1 : 0; o.material = rec.mat; o.t = rec.t;
      o.p[0] = rec.p.x; o.p[1] = rec.p.y; o.p[2] = rec.p.z;
      o.normal[0] = rec.normal.x; o.normal[1] = rec.normal.y; o.normal[2] = rec.normal.z;
      o.u = rec.u; o.v = rec.v;
    } else {
      o.prim = -1; o.material = -1;
      o.t = INF;
    }
  }
  return RTB_OK;
}

int orc_camera_rays(const orc_scene* h, RtbRay* rays) {  // get_ray without the random terms
  const Camera& cam = h->sc.cam;
  for (int j = 0; j < cam.image_height; j++)
    for (int i = 0; i < cam.image_width; i++) {
      Point3 pc = cam.pixel00_loc + ((double)i * cam.pixel_delta_u) + ((double)j * cam.pixel_delta_v);
      Vec3 d = pc - cam.center;
      RtbRay& r = rays[(size_t)j * cam.image_width + i];
      r.origin[0] = cam.center.x; r.origin[1] = cam.center.y; r.origin[2] = cam.center.z;
      r.direction[0] = d.x; r.direction[1] = d.y; r.direction[2] = d.z;
      r.time = 0.; r.t_min = 0.0001;
    }
  return RTB_OK;
}

int orc_medium_interval(orc_scene* h, int32_t medium, const RtbRay* rays, int64_t n, double* t_enter, double* t_exit) {
  const Scene& sc = h->sc;
  int oi = -1;
  for (size_t i = 0; i < sc.objs.size(); i++)
    if (sc.objs[i].kind == RTB_OBJ_MEDIUM && sc.objs[i].medium_index == medium) oi = (int)i;
  if (oi < 0) { g_err = "no such medium"; return RTB_ERR_INVALID; }
  Ctx C;
  C.sc = &sc;
  for (int64_t i = 0; i < n; i++) {
    Ray r(Vec3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]),
          Vec3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), rays[i].time);
    double t1, t2;
    if (medium_interval(C, sc.objs[oi], r, t1, t2)) { t_enter[i] = t1; t_exit[i] = t2; }
    else { t_enter[i] = t_exit[i] = std::numeric_limits<double>::quiet_NaN(); }
  }
  return RTB_OK;
}

int orc_eval_texture(orc_scene* h, int32_t texture, const double* uvp, int64_t n, double* rgb_out) {
  for (int64_t i = 0; i < n; i++) {
    Color c = texture_value(h->sc, texture, uvp[5 * i], uvp[5 * i + 1], Vec3(uvp[5 * i + 2], uvp[5 * i + 3], uvp[5 * i + 4]));
    rgb_out[3 * i] = c.x; rgb_out[3 * i + 1] = c.y; rgb_out[3 * i + 2] = c.z;
  }
  return RTB_OK;
}

int orc_eval_light_pdf(orc_scene* h, const double* od, int64_t n, double* pdf_out) {
  Ctx C;
  C.sc = &h->sc;
  if (h->sc.lights.empty()) { g_err = "scene has no lights"; return RTB_ERR_INVALID; }
  for (int64_t i = 0; i < n; i++)
    pdf_out[i] = lights_pdf_value(C, Vec3(od[6 * i], od[6 * i + 1], od[6 * i + 2]), Vec3(od[6 * i + 3], od[6 * i + 4], od[6 * i + 5]));
  return RTB_OK;
}

// lights.random(origin) n times from one origin (distribution tests of the light sampler)
int orc_sample_lights(orc_scene* h, const double* origin, int64_t n, int sampler_mode, uint64_t stream, double* dir_out) {
  Ctx C;
  C.sc = &h->sc;
  C.S.mode = sampler_mode;
  C.S.key[0] = (uint32_t)h->sc.seed; C.S.key[1] = (uint32_t)(h->sc.seed >> 32);
  C.S.seed_ref(h->sc.seed, stream);
  if (h->sc.lights.empty()) { g_err = "scene has no lights"; return RTB_ERR_INVALID; }
  for (int64_t i = 0; i < n; i++) {
    C.S.set_path((uint32_t)stream, (uint32_t)i);
    C.S.set_bounce(0);
    Vec3 d = lights_random(C, Vec3(origin[0], origin[1], origin[2]));
    dir_out[3 * i] = d.x; dir_out[3 * i + 1] = d.y; dir_out[3 * i + 2] = d.z;
  }
  return RTB_OK;
}

// direction samplers: 0 = random_unit_vector, 1 = random_cosine_direction, 2 = random_in_unit_disk
int orc_sample_directions(int which, int64_t n, int sampler_mode, uint64_t seed, double* out) {
  Sampler S;
  S.mode = sampler_mode;
  S.key[0] = (uint32_t)seed; S.key[1] = (uint32_t)(seed >> 32);
  S.seed_ref(seed, 77);
  for (int64_t i = 0; i < n; i++) {
    S.set_path(7u, (uint32_t)i);
    S.set_bounce(which == 2 ? PRIMARY_BOUNCE : 0u);
    Vec3 d = which == 0 ? random_unit_vector(S) : (which == 1 ? random_cosine_direction(S) : random_in_unit_disk(S));
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
  return RTB_OK;
}

void orc_sphere_uv(const double* p, double* uv) { get_sphere_uv(Vec3(p[0], p[1], p[2]), uv[0], uv[1]); }

void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

// write_color  src/color.rs:8-33 (exposure <= 0: None)
int orc_write_color(const double* pixels_rgb, int64_t n_pixels, double spp, double exposure, uint8_t* rgb8_out) {
  double scale = 1.0 / spp;
  for (int64_t i = 0; i < 3 * n_pixels; i++) {
    double x = pixels_rgb[i] * scale;
    if (exposure > 0.) x = 1. - std::pow(2.718281828459045, -exposure * x);  // color.rs:37-39
    x = linear_to_gamma(x);
    double c = x < 0. ? 0. : (x > 0.999 ? 0.999 : x);  // Interval::clamp; NaN falls through
    rgb8_out[i] = rust_f64_as_u8(256. * c);
  }
  return RTB_OK;
}

// KAT hook: the direction Dielectric::scatter produces (src/material.rs:167-191) for a given uniform draw.
// in = n x {direction[3], face normal[3], front_face, ir, U}
int orc_eval_dielectric(const double* in9, int64_t n, double* dir_out) {
  for (int64_t i = 0; i < n; i++) {
    const double* a = in9 + 9 * i;
    const Vec3 r_in(a[0], a[1], a[2]), normal(a[3], a[4], a[5]);
    const bool front_face = a[6] != 0.;
    const double ir = a[7], U = a[8];
    double refraction_ratio = front_face ? 1.0 / ir : ir;
    Vec3 unit_direction = unit_vector(r_in);
    double cos_theta = std::fmin(dot(-unit_direction, normal), 1.);
    double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    bool cannot_refract = refraction_ratio * sin_theta > 1.0;
    Vec3 direction = (cannot_refract || dielectric_reflectance(cos_theta, refraction_ratio) > U)
                         ? reflect(unit_direction, normal) : refract(unit_direction, normal, refraction_ratio);
    dir_out[3 * i] = direction.x; dir_out[3 * i + 1] = direction.y; dir_out[3 * i + 2] = direction.z;
  }
  return RTB_OK;
}

// auto_expose  src/render.rs:325-339
double orc_auto_expose(const double* pixels_rgb, int64_t n_pixels, double samples_per_pixel) {
  const double medium_weight = 1. / (double)n_pixels;  // 1 / (image_height * image_width)
  double medium_point = 0.;
  for (int64_t i = 0; i < n_pixels; i++) {
    const Color current_color(pixels_rgb[3 * i], pixels_rgb[3 * i + 1], pixels_rgb[3 * i + 2]);
    const double luminance = dot(Color(0.2126, 0.71516, 0.072169), current_color);
    medium_point = medium_point + medium_weight * (luminance * luminance);
  }
  medium_point = medium_point / (samples_per_pixel * samples_per_pixel);
  if (medium_point > 0.001) return -std::log(0.6) / std::sqrt(medium_point);
  return 1.;
}

}  // extern "C"
