// rtb/scene.hpp -- host-side mirror of the reference's scene-construction API.
//
// The reference's host language is Rust; this image has no cargo/rustc, so the host
// side above the C ABI is written in C++ with the same names, argument order and meaning as the
// reference constructors, so that a scene function reads like reference src/main.rs.  (`new` is a
// C++ keyword: constructors are spelled `new_`.)  The source-only Rust crate under rust/ is the
// same thing for a machine that has cargo.
//
// Every type keeps only what the reference constructor receives; the derived fields (bounding
// boxes, quad normal/d/w, sin/cos) are recomputed inside librtb200.so.  `flatten()` walks the
// object tree in `add` order and emits the plain arrays of include/rtb200.h.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/rtb200.h"

namespace rtb {

// ---- vec3.rs ---------------------------------------------------------------------------------
struct Vec3 {
  double x_, y_, z_;
  Vec3() : x_(0), y_(0), z_(0) {}
  Vec3(double x, double y, double z) : x_(x), y_(y), z_(z) {}
  static Vec3 new_(double x, double y, double z) { return Vec3(x, y, z); }  // vec3.rs:29
  static Vec3 new_zero() { return Vec3(); }                                   // vec3.rs:33
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double length_squared() const { return x_ * x_ + y_ * y_ + z_ * z_; }
  double length() const { return std::sqrt(length_squared()); }
};
using Point3 = Vec3;
using Color = Vec3;
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x_ + b.x_, a.y_ + b.y_, a.z_ + b.z_); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x_ - b.x_, a.y_ - b.y_, a.z_ - b.z_); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x_, -a.y_, -a.z_); }
inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a.x_ * b.x_, a.y_ * b.y_, a.z_ * b.z_); }
inline Vec3 operator*(double t, const Vec3& a) { return Vec3(t * a.x_, t * a.y_, t * a.z_); }
inline Vec3 operator*(const Vec3& a, double t) { return t * a; }
inline Vec3 operator/(const Vec3& a, double t) { return Vec3(a.x_ / t, a.y_ / t, a.z_ / t); }  // true division, vec3.rs:145-151
inline Vec3& operator+=(Vec3& a, const Vec3& b) { a = a + b; return a; }   // vec3.rs:74-80 (AddAssign)
inline Vec3& operator*=(Vec3& a, double t) { a = a * t; return a; }        // vec3.rs:121-127 (MulAssign<f64>)
inline Vec3& operator/=(Vec3& a, double t) { a = a / t; return a; }        // vec3.rs:153-157
inline double dot(const Vec3& a, const Vec3& b) { return a.x_ * b.x_ + a.y_ * b.y_ + a.z_ * b.z_; }                 // vec3.rs:167
inline Vec3 cross(const Vec3& a, const Vec3& b) {                                                                    // vec3.rs:171
  return Vec3(a.y_ * b.z_ - a.z_ * b.y_, a.z_ * b.x_ - a.x_ * b.z_, a.x_ * b.y_ - a.y_ * b.x_);
}
inline Vec3 unit_vector(const Vec3& v) { return v / v.length(); }

// ---- utils.rs: host-side construction randomness ----------------------------------------------
// The reference draws scene layouts from the OS-seeded ThreadRng (utils.rs:5-15); here the stream
// is an explicit, seedable xoshiro256++ so that oracle and GPU see the same scene.
class HostRng {
 public:
  explicit HostRng(uint64_t seed = 20240001ull) { reseed(seed); }
  void reseed(uint64_t seed) {
    uint64_t z = seed;
    for (int i = 0; i < 4; i++) {
      z += 0x9E3779B97F4A7C15ull;
      uint64_t t = z;
      t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull;
      t = (t ^ (t >> 27)) * 0x94D049BB133111EBull;
      s_[i] = t ^ (t >> 31);
    }
  }
  uint64_t next() {
    auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
    uint64_t r = rotl(s_[0] + s_[3], 23) + s_[0];
    uint64_t t = s_[1] << 17;
    s_[2] ^= s_[0]; s_[3] ^= s_[1]; s_[1] ^= s_[2]; s_[0] ^= s_[3];
    s_[2] ^= t; s_[3] = rotl(s_[3], 45);
    return r;
  }
 private:
  uint64_t s_[4];
};
inline HostRng& host_rng() { static HostRng r; return r; }
inline void seed_host_rng(uint64_t seed) { host_rng().reseed(seed); }
inline double random_double() { return (double)(host_rng().next() >> 11) * (1.0 / 9007199254740992.0); }
inline double random_range(double lo, double hi) { return lo + (hi - lo) * random_double(); }
inline int64_t random_int(int64_t lo, int64_t hi) { return lo + (int64_t)(host_rng().next() % (uint64_t)(hi - lo + 1)); }
inline Vec3 random_vec3() { double a = random_double(), b = random_double(), c = random_double(); return Vec3(a, b, c); }
inline Vec3 random_vec3_range(double lo, double hi) {
  double a = random_range(lo, hi), b = random_range(lo, hi), c = random_range(lo, hi);
  return Vec3(a, b, c);
}

// ---- texture.rs / perlin.rs / rt_image.rs ------------------------------------------------------
struct TextureNode {
  int kind = RTB_TEX_SOLID;
  Color color;
  double scale = 1.;
  std::shared_ptr<const TextureNode> even, odd;
  int width = 0, height = 0;
  std::shared_ptr<const std::vector<uint8_t>> rgb;  // IMAGE
  std::shared_ptr<const RtbPerlin> perlin;          // NOISE
};
using Texture = std::shared_ptr<const TextureNode>;

struct SolidColor {
  static Texture new_(Color c) {  // texture.rs:33
    auto t = std::make_shared<TextureNode>();
    t->kind = RTB_TEX_SOLID; t->color = c;
    return t;
  }
};
struct CheckerTexture {
  static Texture new_(double scale, Texture even, Texture odd) {  // texture.rs:55 (`_new`)
    auto t = std::make_shared<TextureNode>();
    t->kind = RTB_TEX_CHECKER; t->scale = 1. / scale; t->even = even; t->odd = odd;
    return t;
  }
  static Texture from_color(double scale, Color c1, Color c2) {  // texture.rs:63
    return new_(scale, SolidColor::new_(c1), SolidColor::new_(c2));
  }
};
struct ImageTexture {
  // RGB8 bytes, top row first -- what image::open(..).to_rgb8() returns (rt_image.rs:13-27).
  static Texture from_rgb8(int width, int height, std::vector<uint8_t> bytes) {
    if ((size_t)width * height * 3 != bytes.size()) throw std::invalid_argument("ImageTexture: size mismatch");
    auto t = std::make_shared<TextureNode>();
    t->kind = RTB_TEX_IMAGE; t->width = width; t->height = height;
    t->rgb = std::make_shared<const std::vector<uint8_t>>(std::move(bytes));
    return t;
  }
  // texture.rs:89 takes a file name and decodes it with the `image` crate.  No decoder library is
  // available here; binary PPM (P6, maxval 255) is read, anything else fails like the reference
  // does when the file cannot be opened (rt_image.rs:16-19 panics).
  static Texture new_(const std::string& filename) {
    std::ifstream f(filename, std::ios::binary);
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    if (!(f >> magic >> w >> h >> maxv) || magic != "P6" || maxv != 255 || w <= 0 || h <= 0)
      throw std::runtime_error("Could not open image.");
    f.get();
    std::vector<uint8_t> bytes((size_t)w * h * 3);
    f.read(reinterpret_cast<char*>(bytes.data()), (std::streamsize)bytes.size());
    if (!f) throw std::runtime_error("Could not open image.");
    return from_rgb8(w, h, std::move(bytes));
  }
};
struct Perlin {
  static std::shared_ptr<const RtbPerlin> new_() {  // perlin.rs:15-28, 98-117
    auto p = std::make_shared<RtbPerlin>();
    for (int i = 0; i < 256; i++) {
      Vec3 v = unit_vector(random_vec3_range(-1., 1.));
      p->ranvec[i][0] = v.x(); p->ranvec[i][1] = v.y(); p->ranvec[i][2] = v.z();
    }
    int32_t* perms[3] = {p->perm_x, p->perm_y, p->perm_z};
    for (int k = 0; k < 3; k++) {
      int32_t* a = perms[k];
      for (int i = 0; i < 256; i++) a[i] = i;
      for (int i = 255; i >= 0; i--) {  // permute: (0..n).rev(), target = random_int(0, i)
        int target = (int)random_int(0, i);
        int32_t tmp = a[i]; a[i] = a[target]; a[target] = tmp;
      }
    }
    return p;
  }
};
struct NoiseTexture {
  static Texture new_(double scale) {  // texture.rs:116
    auto t = std::make_shared<TextureNode>();
    t->kind = RTB_TEX_NOISE; t->scale = scale; t->perlin = Perlin::new_();
    return t;
  }
};

// ---- material.rs -------------------------------------------------------------------------------
struct MaterialNode {
  int kind = RTB_MAT_LAMBERTIAN;
  Texture texture;
  Color color;
  double param = 0.;
};
using Material = std::shared_ptr<const MaterialNode>;
inline Material make_material(int kind, Texture t, Color c, double p) {
  auto m = std::make_shared<MaterialNode>();
  m->kind = kind; m->texture = t; m->color = c; m->param = p;
  return m;
}
struct Lambertian {
  static Material new_(Color albedo) { return make_material(RTB_MAT_LAMBERTIAN, SolidColor::new_(albedo), Color(), 0.); }  // :81
  static Material from_texture(Texture t) { return make_material(RTB_MAT_LAMBERTIAN, t, Color(), 0.); }                    // :87
};
struct Metal {
  static Material new_(Color albedo, double f) { return make_material(RTB_MAT_METAL, nullptr, albedo, f < 1. ? f : 1.); }  // :118-121
};
struct Dielectric {
  static Material new_(double ir, Color tint) { return make_material(RTB_MAT_DIELECTRIC, nullptr, tint, ir); }  // :148
  static Material new_clear(double ir) { return new_(ir, Color(1., 1., 1.)); }                                   // :152
};
struct DiffuseLight {
  static Material new_(Color c) { return make_material(RTB_MAT_DIFFUSE_LIGHT, SolidColor::new_(c), Color(), 0.); }  // :200
  static Material from_texture(Texture t) { return make_material(RTB_MAT_DIFFUSE_LIGHT, t, Color(), 0.); }          // :206
};
struct Isotropic {
  static Material new_(Color c) { return make_material(RTB_MAT_ISOTROPIC, SolidColor::new_(c), Color(), 0.); }  // :230
  static Material from_texture(Texture t) { return make_material(RTB_MAT_ISOTROPIC, t, Color(), 0.); }          // :236
};

// ---- object.rs / hittable.rs / transform.rs / constant_medium.rs --------------------------------
struct ObjectNode;
using Object = std::shared_ptr<const ObjectNode>;
struct ObjectNode {
  int kind = RTB_OBJ_LIST;
  Material mat;
  double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<Object> children;  // LIST/BVH: all; TRANSLATE/ROTATE_Y/MEDIUM: exactly one
};

struct Sphere {
  static Object new_(Point3 center, double radius, Material mat) {  // object.rs:83
    auto o = std::make_shared<ObjectNode>();
    o->kind = RTB_OBJ_SPHERE; o->mat = mat;
    o->v[0] = center.x(); o->v[1] = center.y(); o->v[2] = center.z(); o->v[3] = radius;
    return o;
  }
  static Object new_moving(Point3 c1, Point3 c2, double radius, Material mat) {  // object.rs:94
    auto o = std::make_shared<ObjectNode>();
    o->kind = RTB_OBJ_SPHERE; o->mat = mat;
    Vec3 cv = c2 - c1;
    o->v[0] = c1.x(); o->v[1] = c1.y(); o->v[2] = c1.z(); o->v[3] = radius;
    o->v[4] = cv.x(); o->v[5] = cv.y(); o->v[6] = cv.z(); o->v[7] = 1.;
    return o;
  }
};
struct Quad {
  static Object new_(Point3 q, Vec3 u, Vec3 v, Material mat) {  // object.rs:428
    auto o = std::make_shared<ObjectNode>();
    o->kind = RTB_OBJ_QUAD; o->mat = mat;
    o->v[0] = q.x(); o->v[1] = q.y(); o->v[2] = q.z();
    o->v[3] = u.x(); o->v[4] = u.y(); o->v[5] = u.z();
    o->v[6] = v.x(); o->v[7] = v.y(); o->v[8] = v.z();
    return o;
  }
};

class HittableList {  // hittable.rs:55-85
 public:
  std::vector<Object> objects;
  static HittableList new_() { return HittableList(); }
  static HittableList from_object(Object obj) { HittableList l; l.add(obj); return l; }
  void add(Object object) { objects.push_back(object); }
  // create_bvh (hittable.rs:82-84) keeps returning a one-element list wrapping Object::Node; on
  // the device the marker only records that the user asked for acceleration.
  HittableList create_bvh() const {
    auto n = std::make_shared<ObjectNode>();
    n->kind = RTB_OBJ_BVH; n->children = objects;
    return from_object(n);
  }
  Object into_object() const {  // Object::List(Arc::new(list))
    auto n = std::make_shared<ObjectNode>();
    n->kind = RTB_OBJ_LIST; n->children = objects;
    return n;
  }
};
inline Object ObjectList(const HittableList& l) { return l.into_object(); }

inline Object make_box(const Point3& a, const Point3& b, const Material& mat) {  // object.rs:509-560
  HittableList sides;
  Point3 mn(std::fmin(a.x(), b.x()), std::fmin(a.y(), b.y()), std::fmin(a.z(), b.z()));
  Point3 mx(std::fmax(a.x(), b.x()), std::fmax(a.y(), b.y()), std::fmax(a.z(), b.z()));
  Vec3 dx(mx.x() - mn.x(), 0., 0.), dy(0., mx.y() - mn.y(), 0.), dz(0., 0., mx.z() - mn.z());
  sides.add(Quad::new_(Point3(mn.x(), mn.y(), mx.z()), dx, dy, mat));   // front
  sides.add(Quad::new_(Point3(mx.x(), mn.y(), mx.z()), -dz, dy, mat));  // right
  sides.add(Quad::new_(Point3(mx.x(), mn.y(), mn.z()), -dx, dy, mat));  // back
  sides.add(Quad::new_(Point3(mn.x(), mn.y(), mn.z()), dz, dy, mat));   // left
  sides.add(Quad::new_(Point3(mn.x(), mx.y(), mx.z()), dx, -dz, mat));  // top
  sides.add(Quad::new_(Point3(mn.x(), mn.y(), mn.z()), dx, dz, mat));   // bottom
  return sides.into_object();
}

struct Translate {
  static Object new_(Object p, Vec3 displacement) {  // transform.rs:43
    auto o = std::make_shared<ObjectNode>();
    o->kind = RTB_OBJ_TRANSLATE; o->children = {p};
    o->v[0] = displacement.x(); o->v[1] = displacement.y(); o->v[2] = displacement.z();
    return o;
  }
};
struct RotateY {
  static Object new_(Object p, double angle) {  // transform.rs:143
    auto o = std::make_shared<ObjectNode>();
    o->kind = RTB_OBJ_ROTATE_Y; o->children = {p};
    o->v[0] = angle;
    return o;
  }
};
struct ConstantMedium {
  static Object new_(Object boundary, double density, Color c) {  // constant_medium.rs:23
    return from_texture(boundary, density, SolidColor::new_(c));
  }
  static Object from_texture(Object boundary, double density, Texture albedo) {  // :31
    auto o = std::make_shared<ObjectNode>();
    o->kind = RTB_OBJ_MEDIUM; o->children = {boundary}; o->v[0] = density;
    o->mat = Isotropic::from_texture(albedo);
    return o;
  }
};

// object.rs:216-241 -- accepted and ignored by the integrator at HEAD (the sun term of ray_color is commented
// out, render.rs:300-308; Q23).  The records still travel to the library, which adds the term back under
// RTB_FLAG_SUN_LIGHT.
struct Sun {
  Vec3 direction; Color albedo; double limit; double angular_diameter;
  static Sun new_(Vec3 direction, Color albedo, double angular_diameter) {
    return Sun{unit_vector(direction), albedo, 1. - angular_diameter / 180., angular_diameter};
  }
};

// ---- render.rs: Camera ---------------------------------------------------------------------------
struct Camera {  // public fields of render.rs:15-26; the private derived frame lives in the library
  double aspect_ratio; int image_width; int samples_per_pixel; int max_depth; double vfov;
  Point3 lookfrom, lookat; Vec3 vup; double defocus_angle; double focus_dist; Color background;
  bool auto_exposure = false;
  static Camera new_(double aspect_ratio, int image_width, int samples_per_pixel, int max_depth, double vfov,
                     Point3 lookfrom, Point3 lookat, Vec3 vup, double defocus_angle, double focus_dist,
                     Color background) {  // render.rs:62-74
    Camera c;
    c.aspect_ratio = aspect_ratio; c.image_width = image_width; c.samples_per_pixel = samples_per_pixel;
    c.max_depth = max_depth; c.vfov = vfov; c.lookfrom = lookfrom; c.lookat = lookat; c.vup = vup;
    c.defocus_angle = defocus_angle; c.focus_dist = focus_dist; c.background = background;
    return c;
  }
  int image_height() const {  // render.rs:76-77
    int h = (int)((double)image_width / aspect_ratio);
    return h < 1 ? 1 : h;
  }
  RtbCamera to_abi() const {
    RtbCamera a{};
    a.aspect_ratio = aspect_ratio; a.image_width = image_width; a.samples_per_pixel = samples_per_pixel;
    a.max_depth = max_depth; a.vfov = vfov;
    a.lookfrom[0] = lookfrom.x(); a.lookfrom[1] = lookfrom.y(); a.lookfrom[2] = lookfrom.z();
    a.lookat[0] = lookat.x(); a.lookat[1] = lookat.y(); a.lookat[2] = lookat.z();
    a.vup[0] = vup.x(); a.vup[1] = vup.y(); a.vup[2] = vup.z();
    a.defocus_angle = defocus_angle; a.focus_dist = focus_dist;
    a.background[0] = background.x(); a.background[1] = background.y(); a.background[2] = background.z();
    return a;
  }
};

// ---- flatten: object tree -> the arrays of include/rtb200.h ----------------------------------------
class FlatScene {
 public:
  std::vector<RtbObject> objects;
  std::vector<int32_t> children;
  std::vector<int32_t> lights;
  std::vector<RtbMaterial> materials;
  std::vector<RtbTexture> textures;
  std::vector<RtbImage> images;
  std::vector<RtbPerlin> perlins;
  std::vector<std::shared_ptr<const std::vector<uint8_t>>> image_bytes;  // keeps RtbImage.rgb alive
  std::vector<RtbSun> suns;
  int32_t world = -1;
  RtbCamera camera{};
  uint32_t flags = 0;
  uint64_t seed = 20240001ull;

  FlatScene(const Camera& cam, const HittableList& world_list, const HittableList* light_list) {
    camera = cam.to_abi();
    world = emit_list(world_list.objects, RTB_OBJ_LIST);
    if (light_list)
      for (const Object& o : light_list->objects) lights.push_back(emit(o));
  }
  // `lights: Arc<Object>` (render.rs:149): pdf_value / random dispatch on ANY object (object.rs:53-69) -- a list
  // contributes its members, a bare Quad or Sphere is a one-element list, anything else the Hittable defaults
  struct LightObject { Object object; };
  FlatScene(const Camera& cam, const HittableList& world_list, const LightObject& l) {
    camera = cam.to_abi();
    world = emit_list(world_list.objects, RTB_OBJ_LIST);
    if (l.object) {
      if (l.object->kind == RTB_OBJ_LIST) for (const Object& o : l.object->children) lights.push_back(emit(o));
      else lights.push_back(emit(l.object));
    }
  }
  void set_suns(const std::vector<Sun>& in) {
    suns.clear();
    for (const Sun& s : in) {
      RtbSun r{};
      r.direction[0] = s.direction.x(); r.direction[1] = s.direction.y(); r.direction[2] = s.direction.z();
      r.albedo[0] = s.albedo.x(); r.albedo[1] = s.albedo.y(); r.albedo[2] = s.albedo.z();
      r.angular_diameter = s.angular_diameter;
      suns.push_back(r);
    }
  }

  RtbSceneDesc desc() const {
    RtbSceneDesc d{};
    d.abi_version = RTB_ABI_VERSION; d.flags = flags; d.seed = seed;
    d.objects = objects.data(); d.n_objects = (int32_t)objects.size();
    d.children = children.data(); d.n_children = (int32_t)children.size();
    d.world = world;
    d.lights = lights.data(); d.n_lights = (int32_t)lights.size();
    d.materials = materials.data(); d.n_materials = (int32_t)materials.size();
    d.textures = textures.data(); d.n_textures = (int32_t)textures.size();
    d.images = images.data(); d.n_images = (int32_t)images.size();
    d.perlins = perlins.data(); d.n_perlins = (int32_t)perlins.size();
    d.camera = camera;
    d.suns = suns.data(); d.n_suns = (int32_t)suns.size();
    return d;
  }

 private:
  // Shared textures/materials (Arc clones in the reference) are emitted once per distinct node;
  // objects are emitted once per OCCURRENCE so that the graph stays a tree and canonical
  // primitive ids follow `add` order.
  std::vector<std::pair<const void*, int32_t>> tex_seen_, mat_seen_;

  int32_t emit_texture(const Texture& t) {
    if (!t) return -1;
    for (auto& p : tex_seen_) if (p.first == t.get()) return p.second;
    RtbTexture r{};
    r.kind = t->kind; r.a = -1; r.b = -1;
    r.color[0] = t->color.x(); r.color[1] = t->color.y(); r.color[2] = t->color.z();
    r.scale = t->scale;
    if (t->kind == RTB_TEX_CHECKER) { r.a = emit_texture(t->even); r.b = emit_texture(t->odd); }
    if (t->kind == RTB_TEX_IMAGE) {
      RtbImage im{}; im.width = t->width; im.height = t->height; im.rgb = t->rgb->data();
      image_bytes.push_back(t->rgb);
      images.push_back(im);
      r.a = (int32_t)images.size() - 1;
    }
    if (t->kind == RTB_TEX_NOISE) { perlins.push_back(*t->perlin); r.a = (int32_t)perlins.size() - 1; }
    textures.push_back(r);
    int32_t id = (int32_t)textures.size() - 1;
    tex_seen_.push_back({t.get(), id});
    return id;
  }
  int32_t emit_material(const Material& m) {
    if (!m) return -1;
    for (auto& p : mat_seen_) if (p.first == m.get()) return p.second;
    RtbMaterial r{};
    r.kind = m->kind; r.texture = emit_texture(m->texture);
    r.color[0] = m->color.x(); r.color[1] = m->color.y(); r.color[2] = m->color.z();
    r.param = m->param;
    materials.push_back(r);
    int32_t id = (int32_t)materials.size() - 1;
    mat_seen_.push_back({m.get(), id});
    return id;
  }
  int32_t emit_list(const std::vector<Object>& objs, int kind) {
    int32_t self = (int32_t)objects.size();
    objects.push_back(RtbObject{});
    std::vector<int32_t> ids;
    for (const Object& c : objs) ids.push_back(emit(c));
    RtbObject& r = objects[self];
    r.kind = kind; r.material = -1; r.first = (int32_t)children.size(); r.count = (int32_t)ids.size();
    children.insert(children.end(), ids.begin(), ids.end());
    return self;
  }
  int32_t emit(const Object& o) {
    if (o->kind == RTB_OBJ_LIST || o->kind == RTB_OBJ_BVH) return emit_list(o->children, o->kind);
    int32_t self = (int32_t)objects.size();
    objects.push_back(RtbObject{});
    int32_t child = -1;
    if (o->kind == RTB_OBJ_TRANSLATE || o->kind == RTB_OBJ_ROTATE_Y || o->kind == RTB_OBJ_MEDIUM) child = emit(o->children.at(0));
    RtbObject& r = objects[self];
    r.kind = o->kind; r.material = emit_material(o->mat); r.first = child; r.count = child >= 0 ? 1 : 0;
    for (int i = 0; i < 10; i++) r.v[i] = o->v[i];
    return self;
  }
};

}  // namespace rtb
