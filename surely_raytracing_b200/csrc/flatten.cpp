// flatten.cpp -- see flatten.h
#include "flatten.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>

namespace rtb {
namespace {

const double kPi = 3.14159265358979323846;
const double kInf = std::numeric_limits<double>::infinity();

struct D3 { double x, y, z; };
inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double length(D3 a) { return std::sqrt(dot(a, a)); }
inline D3 div(D3 a, double t) { return {a.x / t, a.y / t, a.z / t}; }  // Vec3 / f64 is a true division (src/vec3.rs:145-151)
inline D3 unit(D3 a) { return div(a, length(a)); }

// composed instance transform: p_world = R_y(theta) * p_obj + t, with the matrix convention of
// RotateY::hit's "object -> world" step (src/transform.rs:113-127): x' = c x + s z, z' = -s x + c z
struct Xform {
  double c = 1., s = 0.;
  D3 t = {0., 0., 0.};
  int id = 0;
  D3 rot(D3 p) const { return {c * p.x + s * p.z, p.y, -s * p.x + c * p.z}; }
  D3 point(D3 p) const { return rot(p) + t; }
};

struct Baked {
  int kind, flags, material, xform, id;
  int group = -1, face = 0;  // axis-aligned make_box: group id and face index (2 * axis + side), else -1
  double payload[PRIM_DOUBLES];
  double lo[3], hi[3];
};

struct Builder {
  const RtbSceneDesc& d;
  HostScene& out;
  std::string& err;
  std::vector<char> seen;
  std::vector<Baked> surfaces;
  std::vector<std::vector<Baked>> boundaries;  // per medium
  std::vector<int> medium_material;
  std::vector<double> medium_density;
  int next_id = 0;
  int next_group = 0;
  std::vector<std::array<double, 6>> group_bounds;  // per box group: the exact corner coordinates lo[3], hi[3]

  Builder(const RtbSceneDesc& d_, HostScene& o, std::string& e) : d(d_), out(o), err(e), seen(d_.n_objects, 0) {}

  // Do the six surfaces [start, start + 6) form an axis-aligned box in world space -- make_box (src/object.rs:509-560)
  // under no rotation?  Every quad edge along one axis, every corner coordinate EXACTLY one of the two bounds of its
  // axis (so that a face's plane is the bound itself, bit for bit), every face present once.  Then they become one
  // BVH leaf whose slab test names the face a ray can hit first (rtb_device.cuh, prefilter_box).
  void detect_box(size_t start) {
    double lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
    for (size_t i = start; i < start + 6; i++) {
      const Baked& b = surfaces[i];
      if (b.kind != PRIM_QUAD) return;
      const double* p = b.payload;
      for (int c = 0; c < 4; c++)
        for (int a = 0; a < 3; a++) {
          const double x = p[4 + a] + ((c & 1) ? p[7 + a] : 0.) + ((c & 2) ? p[10 + a] : 0.);
          lo[a] = std::min(lo[a], x); hi[a] = std::max(hi[a], x);
        }
    }
    int faces[6], seen_faces = 0;
    for (size_t i = start; i < start + 6; i++) {
      const double* p = surfaces[i].payload;
      int ua = -1, va = -1;
      for (int a = 0; a < 3; a++) {
        if (p[7 + a] != 0.) { if (ua >= 0) return; ua = a; }
        if (p[10 + a] != 0.) { if (va >= 0) return; va = a; }
      }
      if (ua < 0 || va < 0 || ua == va) return;
      const int k = 3 - ua - va;
      for (int c = 0; c < 4; c++)
        for (int a = 0; a < 3; a++) {
          const double x = (p[4 + a] + ((c & 1) ? p[7 + a] : 0.)) + ((c & 2) ? p[10 + a] : 0.);
          if (x != lo[a] && x != hi[a]) return;
        }
      if (p[4 + k] != lo[k] && p[4 + k] != hi[k]) return;
      // the face must span the other two extents completely
      for (int a : {ua, va}) {
        const double e = a == ua ? p[7 + a] : p[10 + a];
        if (std::min(p[4 + a], p[4 + a] + e) != lo[a] || std::max(p[4 + a], p[4 + a] + e) != hi[a]) return;
      }
      // the unit normal the reference computes must be the exact axis vector (n / |n| with two zero components)
      if (std::fabs(p[k]) != 1. || p[(k + 1) % 3] != 0. || p[(k + 2) % 3] != 0.) return;
      if (p[3] != p[k] * p[4 + k]) return;  // d = normal . q
      const int f = 2 * k + (p[4 + k] == hi[k] ? 1 : 0);
      if (lo[k] == hi[k]) return;
      faces[i - start] = f;
      seen_faces |= 1 << f;
    }
    if (seen_faces != 63) return;
    for (size_t i = start; i < start + 6; i++) { surfaces[i].group = next_group; surfaces[i].face = faces[i - start]; }
    group_bounds.push_back({lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]});
    next_group++;
  }

  bool fail(const std::string& m) { err = m; return false; }

  void bake_quad(const RtbObject& o, const Xform& X, Baked& b) {  // Quad::new  src/object.rs:428-445
    const D3 q = X.point({o.v[0], o.v[1], o.v[2]});
    const D3 u = X.rot({o.v[3], o.v[4], o.v[5]});
    const D3 v = X.rot({o.v[6], o.v[7], o.v[8]});
    const D3 n = cross(u, v);
    const D3 normal = unit(n);
    const D3 w = div(n, dot(n, n));
    const double dd = dot(normal, q);
    const double p16[PRIM_DOUBLES] = {normal.x, normal.y, normal.z, dd, q.x, q.y, q.z, u.x, u.y, u.z, v.x, v.y, v.z, w.x, w.y, w.z};
    std::memcpy(b.payload, p16, sizeof(p16));
    const D3 c[4] = {q, q + u, q + v, q + u + v};
    for (int a = 0; a < 3; a++) { b.lo[a] = kInf; b.hi[a] = -kInf; }
    for (const D3& p : c) {
      const double pc[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], pc[a]); b.hi[a] = std::max(b.hi[a], pc[a]); }
    }
    b.kind = PRIM_QUAD;
    b.flags = 0;
  }

  void bake_sphere(const RtbObject& o, const Xform& X, Baked& b) {  // Sphere::new / new_moving  src/object.rs:83-105
    const D3 c = X.point({o.v[0], o.v[1], o.v[2]});
    const D3 cv = X.rot({o.v[4], o.v[5], o.v[6]});
    const bool moving = o.v[7] != 0.;
    const double r = o.v[3];
    const double p16[PRIM_DOUBLES] = {c.x, c.y, c.z, r, cv.x, cv.y, cv.z, 0., 0., 0., 0., 0., 0., 0., 0., 0.};
    std::memcpy(b.payload, p16, sizeof(p16));
    const double ar = std::fabs(r);
    const double cc[3] = {c.x, c.y, c.z}, vv[3] = {cv.x, cv.y, cv.z};
    for (int a = 0; a < 3; a++) {
      b.lo[a] = cc[a] - ar; b.hi[a] = cc[a] + ar;
      if (moving) { b.lo[a] = std::min(b.lo[a], cc[a] + vv[a] - ar); b.hi[a] = std::max(b.hi[a], cc[a] + vv[a] + ar); }
    }
    b.kind = PRIM_SPHERE;
    b.flags = moving ? PRIM_FLAG_MOVING : 0;
  }

  int xform_id(const Xform& X) {
    for (size_t i = 0; i < out.xforms.size(); i++)
      if (out.xforms[i].x == X.c && out.xforms[i].y == X.s) return (int)i;
    out.xforms.push_back(double2{X.c, X.s});
    return (int)out.xforms.size() - 1;
  }

  // DFS in `add` order; `medium` >= 0 while inside a ConstantMedium boundary
  bool walk(int oi, const Xform& X, int medium, int depth) {
    if (oi < 0 || oi >= d.n_objects) return fail("object index out of range");
    if (depth > 64) return fail("object graph deeper than 64 levels");
    if (seen[oi]) return fail("object referenced twice: the graph must be a tree (flatten emits one node per occurrence)");
    seen[oi] = 1;
    const RtbObject& o = d.objects[oi];
    switch (o.kind) {
      case RTB_OBJ_SPHERE:
      case RTB_OBJ_QUAD: {
        if (o.material < 0 || o.material >= d.n_materials) return fail("material index out of range");
        Baked b;
        if (o.kind == RTB_OBJ_QUAD) bake_quad(o, X, b);
        else bake_sphere(o, X, b);
        b.material = o.material;
        b.xform = xform_id(X);
        b.id = next_id++;
        for (int a = 0; a < PRIM_DOUBLES; a++)
          if (!std::isfinite(b.payload[a])) return fail("non-finite primitive (degenerate quad or bad coordinates)");
        (medium >= 0 ? boundaries[medium] : surfaces).push_back(b);
        return true;
      }
      case RTB_OBJ_LIST:
      case RTB_OBJ_BVH: {
        if (o.first < 0 || o.count < 0 || (long long)o.first + o.count > d.n_children) return fail("list child range out of bounds");
        const size_t start = surfaces.size();
        for (int k = 0; k < o.count; k++)
          if (!walk(d.children[o.first + k], X, medium, depth + 1)) return false;
        if (medium < 0 && o.count == 6 && surfaces.size() == start + 6 && !(d.flags & (RTB_FLAG_BVH_LEAF4 | RTB_FLAG_NO_BOX_LEAVES))) detect_box(start);
        return true;
      }
      case RTB_OBJ_TRANSLATE: {  // src/transform.rs:57-69
        Xform Y = X;
        Y.t = X.t + X.rot({o.v[0], o.v[1], o.v[2]});
        return walk(o.first, Y, medium, depth + 1);
      }
      case RTB_OBJ_ROTATE_Y: {  // src/transform.rs:143-146
        const double radians = o.v[0] * (kPi / 180.);
        const double sa = std::sin(radians), ca = std::cos(radians);
        Xform Y = X;
        Y.c = X.c * ca - X.s * sa;
        Y.s = X.s * ca + X.c * sa;
        return walk(o.first, Y, medium, depth + 1);
      }
      case RTB_OBJ_MEDIUM: {  // src/constant_medium.rs:23-29
        if (medium >= 0) { err = "a ConstantMedium used as the boundary of another medium is not supported"; return false; }
        if (o.material < 0 || o.material >= d.n_materials) return fail("medium material index out of range");
        if (!(o.v[0] > 0.)) return fail("medium density must be positive");
        boundaries.emplace_back();
        medium_material.push_back(o.material);
        medium_density.push_back(o.v[0]);
        return walk(o.first, X, (int)boundaries.size() - 1, depth + 1);
      }
      default:
        return fail("unknown object kind");
    }
  }
};

// ---- SAH BVH2 over surface primitives -------------------------------------------------------------
struct Box {
  double lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
  void grow(const double* l, const double* h) {
    for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); }
  }
  double area() const {
    const double x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    if (x < 0. || y < 0. || z < 0.) return 0.;
    return 2. * (x * y + y * z + z * x);
  }
};

struct BuildNode {
  Box box;
  int left = -1, right = -1;   // children (BuildNode indices); -1 for a leaf
  int first = 0, count = 0;    // leaf range in the ordered primitive index array
};

struct BuildBounds { double lo[3], hi[3]; };  // what the builder reads of an item, packed (the Baked records are 4x larger: cache misses)

struct BvhBuilder {
  std::vector<BuildBounds> prims;
  std::vector<int> order;
  std::vector<BuildNode> nodes;
  int max_depth = 0;
  std::vector<std::pair<double, int>> keyed;  // scratch of the exact sweep
  std::vector<double> sweep_area;
  std::vector<int> part_scratch;              // right-hand side of the stable partition
  double bin_cmin[3], bin_cmax[3];            // centroid range of the node being split (binned search -> partition)
  double kTraversal = 1.0, kIntersect = 3.0;  // SAH costs
  int max_leaf = 1;  // measured on c4: one primitive per leaf, Ci/Ct = 3 -> fewest f64 tests (profiles/r01_bvh_sweep.txt)

  BvhBuilder(const std::vector<Baked>& p, bool leaves_of_4) : prims(p.size()), order(p.size()) {
    for (size_t i = 0; i < p.size(); i++)
      for (int a = 0; a < 3; a++) { prims[i].lo[a] = p[i].lo[a]; prims[i].hi[a] = p[i].hi[a]; }
    std::iota(order.begin(), order.end(), 0);
    if (leaves_of_4) { max_leaf = 4; kIntersect = 0.7; }  // RTB_FLAG_BVH_LEAF4: the tuning arm of profiles/r01_bvh_sweep.txt
  }

  int build(int first, int count, int depth) {
    max_depth = std::max(max_depth, depth);
    BuildNode node;
    for (int i = 0; i < count; i++) node.box.grow(prims[order[first + i]].lo, prims[order[first + i]].hi);
    node.first = first;
    node.count = count;
    const int self = (int)nodes.size();
    nodes.push_back(node);
    if (count <= 1) return self;
    // SAH split search: exact sweep over the sorted centroids for small nodes, 32 centroid bins above
    // (scene_create sits on the e2e path of every render_par call, so the build must stay ~ms).
    double best_cost = kInf;
    int best_axis = -1, best_split = -1;
    double best_plane = 0.;  // binned split: centroid*2 threshold on best_axis
    const bool binned = count > 8;
    if (!binned) {
      if ((int)keyed.size() < count) { keyed.resize(count); sweep_area.resize(count); }
      for (int axis = 0; axis < 3; axis++) {
        for (int i = 0; i < count; i++) {
          const int pi = order[first + i];
          keyed[i] = {prims[pi].lo[axis] + prims[pi].hi[axis], i};  // (key, position): ties keep their order
        }
        std::sort(keyed.begin(), keyed.begin() + count);
        Box rb;
        for (int i = count - 1; i > 0; i--) {
          const BuildBounds& p = prims[order[first + keyed[i].second]];
          rb.grow(p.lo, p.hi);
          sweep_area[i] = rb.area();
        }
        Box lb;
        for (int i = 1; i < count; i++) {
          const BuildBounds& p = prims[order[first + keyed[i - 1].second]];
          lb.grow(p.lo, p.hi);
          const double cost = lb.area() * i + sweep_area[i] * (count - i);
          if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i; }
        }
      }
    } else {
      // one pass for the centroid range of all three axes, one pass for all three binnings, then a sweep per axis
      constexpr int NB = 32;
      for (int a = 0; a < 3; a++) { bin_cmin[a] = kInf; bin_cmax[a] = -kInf; }
      for (int i = 0; i < count; i++) {
        const BuildBounds& p = prims[order[first + i]];
        for (int a = 0; a < 3; a++) {
          const double c = p.lo[a] + p.hi[a];
          bin_cmin[a] = std::min(bin_cmin[a], c); bin_cmax[a] = std::max(bin_cmax[a], c);
        }
      }
      Box bins[3][NB];
      int cnt[3][NB] = {};
      double scale3[3];
      for (int a = 0; a < 3; a++) scale3[a] = bin_cmax[a] > bin_cmin[a] ? NB / (bin_cmax[a] - bin_cmin[a]) : 0.;
      for (int i = 0; i < count; i++) {
        const BuildBounds& p = prims[order[first + i]];
        for (int a = 0; a < 3; a++) {
          if (scale3[a] == 0.) continue;
          int bi = (int)((p.lo[a] + p.hi[a] - bin_cmin[a]) * scale3[a]);
          bi = bi < 0 ? 0 : (bi >= NB ? NB - 1 : bi);
          bins[a][bi].grow(p.lo, p.hi);
          cnt[a][bi]++;
        }
      }
      for (int axis = 0; axis < 3; axis++) {
        if (scale3[axis] == 0.) continue;
        const double cmin = bin_cmin[axis], scale = scale3[axis];
        double right_area[NB];
        int right_cnt[NB];
        Box rb;
        int rc = 0;
        for (int b = NB - 1; b > 0; b--) { if (cnt[axis][b]) rb.grow(bins[axis][b].lo, bins[axis][b].hi); rc += cnt[axis][b]; right_area[b] = rb.area(); right_cnt[b] = rc; }
        Box lb;
        int lc = 0;
        for (int b = 1; b < NB; b++) {
          if (cnt[axis][b - 1]) lb.grow(bins[axis][b - 1].lo, bins[axis][b - 1].hi);
          lc += cnt[axis][b - 1];
          if (lc == 0 || right_cnt[b] == 0) continue;
          const double cost = lb.area() * lc + right_area[b] * right_cnt[b];
          if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = lc; best_plane = cmin + b / scale; }
        }
      }
      if (best_axis < 0) {  // all centroids coincide: split by index
        best_axis = 0; best_split = count / 2; best_cost = node.box.area() * count;
      }
    }
    const double parent_area = std::max(node.box.area(), 1e-300);
    const double split_cost = kTraversal + kIntersect * best_cost / parent_area;
    const double leaf_cost = kIntersect * count;
    if (count <= max_leaf && leaf_cost <= split_cost) return self;
    if (binned && best_plane != 0.) {
      // same bin assignment as the search (so the counts match), order inside each side kept
      const double cmin = bin_cmin[best_axis], cmax = bin_cmax[best_axis];
      const double scale = 32 / (cmax - cmin);
      const int split_bin = (int)std::lround((best_plane - cmin) * scale);
      if ((int)part_scratch.size() < count) part_scratch.resize(count);
      int nl = 0, nr = 0;
      for (int i = 0; i < count; i++) {
        const int a = order[first + i];
        int b = (int)((prims[a].lo[best_axis] + prims[a].hi[best_axis] - cmin) * scale);
        b = b < 0 ? 0 : (b >= 32 ? 31 : b);
        if (b < split_bin) order[first + nl++] = a;  // (nl <= i: never overwrites an unread entry)
        else part_scratch[nr++] = a;
      }
      std::copy(part_scratch.begin(), part_scratch.begin() + nr, order.begin() + first + nl);
      best_split = nl;
      if (best_split <= 0 || best_split >= count) {  // numerical corner: fall back to a median split
        std::stable_sort(order.begin() + first, order.begin() + first + count, [&](int a, int b) {
          return prims[a].lo[best_axis] + prims[a].hi[best_axis] < prims[b].lo[best_axis] + prims[b].hi[best_axis];
        });
        best_split = count / 2;
      }
    } else {
      std::stable_sort(order.begin() + first, order.begin() + first + count, [&](int a, int b) {
        return prims[a].lo[best_axis] + prims[a].hi[best_axis] < prims[b].lo[best_axis] + prims[b].hi[best_axis];
      });
    }
    const int l = build(first, best_split, depth + 1);
    const int r = build(first + best_split, count - best_split, depth + 1);
    nodes[self].left = l;
    nodes[self].right = r;
    nodes[self].count = 0;
    return self;
  }
};

inline float round_down(double x) { float f = (float)x; return ((double)f > x) ? std::nextafterf(f, -INFINITY) : f; }
inline float round_up(double x) { float f = (float)x; return ((double)f < x) ? std::nextafterf(f, INFINITY) : f; }
// the two fp32 slots of one axis of a node box: outward-rounded lo / hi (default) or, in the
// RTB_SLAB_CENTER A/B arm, centre / half-extent of a box that still contains [lo, hi]
inline void box_center_half(double lo, double hi, float& c, float& h) {
#if !defined(RTB_SLAB_CENTER)
  c = round_down(lo); h = round_up(hi);
  return;
#endif
  c = (float)(0.5 * (lo + hi));
  h = round_up(std::max((double)c - lo, hi - (double)c));
  h = std::nextafterf(h, INFINITY);
}

int shading_class(const RtbSceneDesc& d, int material) {
  const RtbMaterial& m = d.materials[material];
  switch (m.kind) {
    case RTB_MAT_DIFFUSE_LIGHT: return CLS_LIGHT;
    case RTB_MAT_METAL: return CLS_METAL;
    case RTB_MAT_DIELECTRIC: return CLS_DIELECTRIC;
    case RTB_MAT_ISOTROPIC: return CLS_ISOTROPIC;
    default: {
      const int tk = d.textures[m.texture].kind;
      return tk == RTB_TEX_SOLID ? CLS_LAMBERT_SOLID : (tk == RTB_TEX_NOISE ? CLS_NOISE : CLS_LAMBERT_TEX);
    }
  }
}

void emit_prim(const RtbSceneDesc& d, HostScene& out, const Baked& b) {
  out.prims.insert(out.prims.end(), b.payload, b.payload + PRIM_DOUBLES);
  DPre pre{};
  if (b.kind == PRIM_QUAD) {  // payload: n d | q | u | v | w
    const double* p = b.payload;
    const D3 u = {p[7], p[8], p[9]}, v = {p[10], p[11], p[12]}, w = {p[13], p[14], p[15]};
    const D3 A = cross(v, w), B = cross(w, u);
    pre.qx = p[4]; pre.qy = p[5]; pre.qz = p[6];
    pre.ax = (float)A.x; pre.ay = (float)A.y; pre.az = (float)A.z;
    pre.bx = (float)B.x; pre.by = (float)B.y; pre.bz = (float)B.z;
    pre.nx = (float)p[0]; pre.ny = (float)p[1]; pre.nz = (float)p[2];
    const double a1 = std::fabs(A.x) + std::fabs(A.y) + std::fabs(A.z), b1 = std::fabs(B.x) + std::fabs(B.y) + std::fabs(B.z);
    pre.ab1 = std::nextafterf(round_up(std::max(a1, b1)), INFINITY);
  }
  out.pre.push_back(pre);
  const int mat_bits = (b.material >= 0 && b.material <= PRIM_MAT_MAX) ? ((b.material + 1) << PRIM_MAT_SHIFT) : 0;
  out.prim_info.push_back(int4{b.kind | b.flags | (shading_class(d, b.material) << PRIM_CLASS_SHIFT) | mat_bits, b.material, b.xform, b.id});
}

bool texture_needs_uv(const RtbSceneDesc& d, int ti, int depth = 0) {
  if (ti < 0 || ti >= d.n_textures || depth > 16) return false;
  const RtbTexture& t = d.textures[ti];
  if (t.kind == RTB_TEX_IMAGE) return true;
  if (t.kind == RTB_TEX_CHECKER) return texture_needs_uv(d, t.a, depth + 1) || texture_needs_uv(d, t.b, depth + 1);
  return false;
}

}  // namespace

int flatten_scene(const RtbSceneDesc& d, HostScene& out, std::string& err) {
  if (d.abi_version != RTB_ABI_VERSION) { err = "abi version mismatch"; return RTB_ERR_INVALID; }
  if (d.n_objects <= 0 || !d.objects) { err = "empty object array"; return RTB_ERR_INVALID; }
  if (d.world < 0 || d.world >= d.n_objects ||
      (d.objects[d.world].kind != RTB_OBJ_LIST && d.objects[d.world].kind != RTB_OBJ_BVH)) {
    err = "world must be a list object";
    return RTB_ERR_INVALID;
  }
  out = HostScene();
  out.flags = d.flags;
  out.seed = d.seed;
  out.xforms.push_back(double2{1., 0.});  // id 0 = identity

  // ---- textures / materials -----------------------------------------------------------------
  for (int i = 0; i < d.n_textures; i++) {
    const RtbTexture& t = d.textures[i];
    DTexture o{};
    o.kind = t.kind; o.a = t.a; o.b = t.b;
    o.color[0] = (float)t.color[0]; o.color[1] = (float)t.color[1]; o.color[2] = (float)t.color[2];
    o.scale = t.scale;
    switch (t.kind) {
      case RTB_TEX_SOLID: break;
      case RTB_TEX_CHECKER:
        if (t.a < 0 || t.a >= d.n_textures || t.b < 0 || t.b >= d.n_textures) { err = "checker child texture out of range"; return RTB_ERR_INVALID; }
        break;
      case RTB_TEX_IMAGE: {
        if (t.a < 0 || t.a >= d.n_images) { err = "image index out of range"; return RTB_ERR_INVALID; }
        const RtbImage& im = d.images[t.a];
        if (im.width <= 0 || im.height < 0 || (im.height > 0 && !im.rgb)) { err = "bad image"; return RTB_ERR_INVALID; }
        o.width = im.width; o.height = im.height;
        o.a = (int)out.texels.size();
        out.texels.insert(out.texels.end(), im.rgb, im.rgb + (size_t)3 * im.width * im.height);
        break;
      }
      case RTB_TEX_NOISE:
        if (t.a < 0 || t.a >= d.n_perlins) { err = "perlin index out of range"; return RTB_ERR_INVALID; }
        break;
      default: err = "unknown texture kind"; return RTB_ERR_INVALID;
    }
    out.textures.push_back(o);
  }
  for (int i = 0; i < d.n_perlins; i++) {
    const RtbPerlin& p = d.perlins[i];
    for (int k = 0; k < 256; k++) out.perlin_vec.push_back(float4{(float)p.ranvec[k][0], (float)p.ranvec[k][1], (float)p.ranvec[k][2], 0.f});
    const int32_t* perms[3] = {p.perm_x, p.perm_y, p.perm_z};
    for (int a = 0; a < 3; a++)
      for (int k = 0; k < 256; k++) {
        if (perms[a][k] < 0 || perms[a][k] > 255) { err = "perlin permutation entry out of range"; return RTB_ERR_INVALID; }
        out.perlin_perm.push_back((uint8_t)perms[a][k]);
      }
  }
  for (int i = 0; i < d.n_materials; i++) {
    const RtbMaterial& m = d.materials[i];
    DMaterial o{};
    o.kind = m.kind; o.texture = m.texture;
    o.color[0] = (float)m.color[0]; o.color[1] = (float)m.color[1]; o.color[2] = (float)m.color[2];
    o.param = (float)m.param;
    if (m.kind < RTB_MAT_LAMBERTIAN || m.kind > RTB_MAT_ISOTROPIC) { err = "unknown material kind"; return RTB_ERR_INVALID; }
    const bool textured = m.kind == RTB_MAT_LAMBERTIAN || m.kind == RTB_MAT_DIFFUSE_LIGHT || m.kind == RTB_MAT_ISOTROPIC;
    if (textured && (m.texture < 0 || m.texture >= d.n_textures)) { err = "material texture index out of range"; return RTB_ERR_INVALID; }
    o.needs_uv = textured && texture_needs_uv(d, m.texture) ? 1 : 0;
    out.materials.push_back(o);
  }

  // ---- object graph ---------------------------------------------------------------------------
  Builder B(d, out, err);
  if (!B.walk(d.world, Xform(), -1, 0)) return err.find("not supported") != std::string::npos ? RTB_ERR_UNSUPPORTED : RTB_ERR_INVALID;
  if (B.surfaces.empty() && B.boundaries.empty()) { /* an empty world is legal: every ray misses */ }

  // ---- lights: the `lights` list (src/main.rs:485-494); only Quad / Sphere sample, others are
  //      the Hittable defaults pdf 0 / direction (1,0,0) (src/hittable.rs:46-52, object.rs:53-69)
  for (int k = 0; k < d.n_lights; k++) {
    const int li = d.lights[k];
    if (li < 0 || li >= d.n_objects) { err = "light index out of range"; return RTB_ERR_INVALID; }
    const RtbObject& o = d.objects[li];
    DLight L{};
    Baked b{};
    if (o.kind == RTB_OBJ_QUAD) {
      B.bake_quad(o, Xform(), b);
      L.kind = LIGHT_QUAD;
      const D3 u = {o.v[3], o.v[4], o.v[5]}, v = {o.v[6], o.v[7], o.v[8]};
      L.area = length(cross(u, v));
      for (int a = 0; a < 3; a++) { L.q[a] = o.v[a]; L.u[a] = o.v[3 + a]; L.v[a] = o.v[6 + a]; }
    } else if (o.kind == RTB_OBJ_SPHERE) {
      B.bake_sphere(o, Xform(), b);
      L.kind = LIGHT_SPHERE;
    } else if (o.kind == RTB_OBJ_LIST || o.kind == RTB_OBJ_BVH) {
      err = "nested lists inside the light list are not supported";
      return RTB_ERR_UNSUPPORTED;
    } else {
      L.kind = LIGHT_OTHER;
    }
    std::memcpy(L.prim, b.payload, sizeof(L.prim));
    out.lights.push_back(L);
  }

  // ---- suns: accepted and ignored at HEAD (Q23); RTB_FLAG_SUN_LIGHT switches the commented-out term back on
  if (d.n_suns < 0 || (d.n_suns > 0 && !d.suns)) { err = "bad sun array"; return RTB_ERR_INVALID; }
  if (d.flags & RTB_FLAG_SUN_LIGHT) {
    if (d.n_suns > MAX_SUNS) { err = "more suns than the backend holds (4)"; return RTB_ERR_UNSUPPORTED; }
    for (int k = 0; k < d.n_suns; k++) {  // Sun::new  src/object.rs:223-231
      const D3 dir = unit({d.suns[k].direction[0], d.suns[k].direction[1], d.suns[k].direction[2]});
      if (!std::isfinite(dir.x + dir.y + dir.z)) { err = "degenerate sun direction"; return RTB_ERR_INVALID; }
      DSun& o = out.suns[out.n_suns++];
      o.dir[0] = (float)dir.x; o.dir[1] = (float)dir.y; o.dir[2] = (float)dir.z;
      for (int a = 0; a < 3; a++) o.albedo[a] = (float)d.suns[k].albedo[a];
      o.limit = (float)(1. - d.suns[k].angular_diameter / 180.);
      o.pad = 0.f;
    }
  }

  // ---- camera: Camera::new  src/render.rs:62-134 ------------------------------------------------
  {
    const RtbCamera& c = d.camera;
    if (c.image_width <= 0 || c.samples_per_pixel <= 0 || c.max_depth < 0 || !(c.aspect_ratio > 0.)) { err = "bad camera"; return RTB_ERR_INVALID; }
    int image_height = (int)((double)c.image_width / c.aspect_ratio);
    if (image_height < 1) image_height = 1;
    const D3 lookfrom = {c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]}, lookat = {c.lookat[0], c.lookat[1], c.lookat[2]};
    const D3 vup = {c.vup[0], c.vup[1], c.vup[2]};
    const double theta = c.vfov * (kPi / 180.);
    const double h = std::tan(theta / 2.);
    const double focus_dist = c.focus_dist <= 0. ? 1. : c.focus_dist;  // Q2
    const double viewport_height = 2. * h * focus_dist;
    const double viewport_width = viewport_height * (double)c.image_width / (double)image_height;
    const D3 w = unit(lookfrom - lookat), u = unit(cross(vup, w)), v = cross(w, u);
    const D3 viewport_u = viewport_width * u, viewport_v = viewport_height * ((-1.) * v);
    const D3 du = div(viewport_u, (double)c.image_width), dv = div(viewport_v, (double)image_height);
    const D3 upper_left = lookfrom - (focus_dist * w) - div(viewport_u, 2.) - div(viewport_v, 2.);
    const D3 pixel00 = upper_left + 0.5 * (du + dv);
    const double defocus_radius = focus_dist * std::tan((c.defocus_angle / 2.) * (kPi / 180.));
    const int root = (int)std::sqrt((double)c.samples_per_pixel);
    const int spp = root * root;  // nearest_square (Q1)
    const double sqrt_spp = std::sqrt((double)spp);
    DCamera& C = out.cam;
    std::memset(&C, 0, sizeof(C));
    const D3 disk_u = defocus_radius * u, disk_v = defocus_radius * v;
    const D3 vs[6] = {lookfrom, pixel00, du, dv, disk_u, disk_v};
    double* dst[6] = {C.center, C.pixel00, C.du, C.dv, C.disk_u, C.disk_v};
    for (int k = 0; k < 6; k++) { dst[k][0] = vs[k].x; dst[k][1] = vs[k].y; dst[k][2] = vs[k].z; }
    C.recip_sqrt_spp = 1. / sqrt_spp;
    C.width = c.image_width; C.height = image_height; C.sqrt_spp = (int)sqrt_spp; C.spp = spp;
    C.max_depth = c.max_depth; C.defocus = c.defocus_angle <= 0. ? 0 : 1;
    for (int a = 0; a < 3; a++) C.background[a] = (float)c.background[a];
    if (spp < 1) { err = "samples_per_pixel rounds down to zero"; return RTB_ERR_INVALID; }
    if ((long long)C.width * C.height > (1ll << 31) - 1) { err = "image too large"; return RTB_ERR_INVALID; }
  }

  // ---- BVH over the surfaces ----------------------------------------------------------------------
  // fp32 slab padding: covers the rounding of the f64 origin to fp32 (<= 2^-24 |o|) for origins up to
  // ~8x the scene magnitude M; the multiplicative slack in the kernel covers the slab arithmetic.
  double M = 1.;
  for (const Baked& b : B.surfaces)
    for (int a = 0; a < 3; a++) M = std::max(M, std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])));
  for (const auto& bl : B.boundaries)
    for (const Baked& b : bl)
      for (int a = 0; a < 3; a++) M = std::max(M, std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])));
  for (int a = 0; a < 3; a++) M = std::max(M, std::fabs(out.cam.center[a]));
  const double pad = M * 1e-6;
  {
    double mag = M;
    for (int a = 0; a < 3; a++) mag = std::max(mag, std::fabs(out.cam.center[a]) + std::fabs(out.cam.disk_u[a]) + std::fabs(out.cam.disk_v[a]));
    out.scene_mag = round_up(mag);
  }

  // The BVH is built over ITEMS: a surface primitive, or an axis-aligned box group (six quads: one leaf).
  std::vector<Baked> items;
  std::vector<std::vector<int>> members;
  {
    std::vector<int> group_item(B.next_group, -1);
    for (size_t i = 0; i < B.surfaces.size(); i++) {
      const Baked& b = B.surfaces[i];
      if (b.group < 0) { items.push_back(b); members.push_back({(int)i}); continue; }
      if (group_item[b.group] < 0) {
        group_item[b.group] = (int)items.size();
        items.push_back(b);
        members.push_back(std::vector<int>(6, -1));
      }
      Baked& it = items[group_item[b.group]];
      for (int a = 0; a < 3; a++) { it.lo[a] = std::min(it.lo[a], b.lo[a]); it.hi[a] = std::max(it.hi[a], b.hi[a]); }
      members[group_item[b.group]][b.face] = (int)i;
    }
  }
  BvhBuilder bvh(items, (d.flags & RTB_FLAG_BVH_LEAF4) != 0);
  std::vector<int> node_remap;
  if (!items.empty()) bvh.build(0, (int)items.size(), 0);
  if (bvh.max_depth + 2 > BVH_STACK) { err = "BVH deeper than the traversal stack"; return RTB_ERR_UNSUPPORTED; }
  if (B.surfaces.size() >= (size_t)1 << 26) { err = "more than 2^26 surface primitives (leaf references hold 26 index bits)"; return RTB_ERR_UNSUPPORTED; }
  out.bvh_depth = bvh.max_depth;
  out.multi_leaf = 0;
  for (const BuildNode& bn : bvh.nodes) out.multi_leaf |= (bn.left < 0 && bn.count > 1) ? 1 : 0;
  std::vector<int> item_first(items.size(), 0);  // emitted index of the first primitive of the item at each position of bvh.order
  for (size_t pos = 0; pos < bvh.order.size(); pos++) {
    item_first[pos] = (int)out.prim_info.size();
    const std::vector<int>& mem = members[bvh.order[pos]];
    for (int m : mem) emit_prim(d, out, B.surfaces[m]);
    if (mem.size() == 6 && items[bvh.order[pos]].group >= 0) {  // the box bounds go into the DPre slot of the first face
      DBoxBounds bb{};
      const Baked& it = items[bvh.order[pos]];
      for (int a = 0; a < 3; a++) { bb.lo[a] = B.group_bounds[it.group][a]; bb.hi[a] = B.group_bounds[it.group][3 + a]; }  // (not the padded AABB)
      static_assert(sizeof(DBoxBounds) == sizeof(DPre), "");
      std::memcpy(&out.pre[item_first[pos]], &bb, sizeof(bb));
    }
  }
  out.n_surface_prims = (int)B.surfaces.size();

  // Quantisation grid of the 32-byte nodes: 16-bit cell indices over the padded root box, per axis.
  // Child boxes are rounded outward and then widened by QPAD cells: the kernel evaluates
  // t = (2^23 + k) * inv_d - (2^23 + o') * inv_d with the second product rounded to fp32, which moves
  // every plane by at most 0.5 (1 + |o'| / 2^23) <= 1 cell for ray origins within 2^23 cells of the
  // grid (beyond that the kernel traverses unculled).
  constexpr int QPAD = 2;
  if (!B.surfaces.empty()) {
    const Box& rb = bvh.nodes[0].box;
    for (int a = 0; a < 3; a++) {
      const double lo = rb.lo[a] - pad, hi = rb.hi[a] + pad;
      const double cell = std::max(hi - lo, 1e-9 * M) / 65500.;
      out.grid_cell[a] = (float)cell;
      out.grid_base[a] = lo - 16. * (double)out.grid_cell[a];
      out.grid_inv_cell[a] = 1. / (double)out.grid_cell[a];
    }
  }
  auto quant = [&](const Box& bx, int a) -> unsigned {
    const double lo = (bx.lo[a] - pad - out.grid_base[a]) * out.grid_inv_cell[a];
    const double hi = (bx.hi[a] + pad - out.grid_base[a]) * out.grid_inv_cell[a];
    const long long klo = std::max<long long>(0, (long long)std::floor(lo) - QPAD);
    const long long khi = std::min<long long>(65535, (long long)std::ceil(hi) + QPAD);
    return (unsigned)klo | ((unsigned)khi << 16);
  };
  // expected-visit inflation of the grid boxes: sum of surface areas, quantised vs exact (SAH measure)
  double sa_exact = 0., sa_quant = 0.;
  auto emit_qnode = [&](int self, const Box& b0, const Box& b1, int ref0, int ref1) {
    if (out.qnodes.size() < 2 * (size_t)(self + 1)) out.qnodes.resize(2 * (size_t)(self + 1));
    const Box* bs[2] = {&b0, &b1};
    const int refs2[2] = {ref0, ref1};
    for (int c = 0; c < 2; c++) {
      const unsigned q[3] = {quant(*bs[c], 0), quant(*bs[c], 1), quant(*bs[c], 2)};
      out.qnodes[2 * (size_t)self + c] = uint4{q[0], q[1], q[2], (unsigned)refs2[c]};
      double e[3], g[3];
      for (int a = 0; a < 3; a++) {
        e[a] = bs[c]->hi[a] - bs[c]->lo[a] + 2. * pad;
        g[a] = (double)((q[a] >> 16) - (q[a] & 0xFFFFu)) * (double)out.grid_cell[a];
      }
      sa_exact += e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
      sa_quant += g[0] * g[1] + g[1] * g[2] + g[2] * g[0];
    }
  };

  // inner nodes get consecutive device indices in DFS order (root = 0)
  auto leaf_ref = [&](int first, int count) {  // `first`, `count`: positions of bvh.order (items)
    const Baked& b0 = items[bvh.order[first]];
    if (b0.group >= 0) return leaf_make(item_first[first], 6, LEAF_KIND_BOX);  // (box groups only exist with one-item leaves)
    const int bits = (b0.kind == PRIM_QUAD ? LEAF_KIND_QUAD : 0) | ((b0.flags & PRIM_FLAG_MOVING) ? LEAF_KIND_MOVING : 0);
    return leaf_make(item_first[first], count, bits);
  };
  // Device indices of the inner nodes: the top TOP_LEVELS levels breadth-first (so that "the first k nodes" are the
  // top of the tree: what the shared-memory arm of the extend kernel stages), the subtrees below them depth-first
  // (a ray's consecutive visits stay close in memory).
  constexpr int TOP_LEVELS = 7;
  std::vector<int> dev_index(bvh.nodes.size(), -1);
  {
    int next = 0;
    std::vector<std::pair<int, int>> frontier, below;  // (build node, depth)
    if (!bvh.nodes.empty() && bvh.nodes[0].left >= 0) frontier.push_back({0, 0});
    for (size_t k = 0; k < frontier.size(); k++) {
      const int bi = frontier[k].first, depth = frontier[k].second;
      dev_index[bi] = next++;
      for (int child : {bvh.nodes[bi].left, bvh.nodes[bi].right})
        if (bvh.nodes[child].left >= 0) (depth + 1 < TOP_LEVELS ? frontier : below).push_back({child, depth + 1});
    }
    std::function<void(int)> dfs = [&](int bi) {
      dev_index[bi] = next++;
      for (int child : {bvh.nodes[bi].left, bvh.nodes[bi].right})
        if (bvh.nodes[child].left >= 0) dfs(child);
    };
    for (const auto& b : below) dfs(b.first);
    out.nodes.resize(4 * (size_t)next);
  }
  std::function<int(int)> emit_node = [&](int bi) -> int {
    const BuildNode& bn = bvh.nodes[bi];
    if (bn.left < 0) return leaf_ref(bn.first, bn.count);
    const int self = dev_index[bi];
    const int refs[2] = {emit_node(bn.left), emit_node(bn.right)};
    const Box* cb[2] = {&bvh.nodes[bn.left].box, &bvh.nodes[bn.right].box};
    float ctr[2][3], half[2][3];
    for (int c = 0; c < 2; c++)
      for (int a = 0; a < 3; a++) box_center_half(cb[c]->lo[a] - pad, cb[c]->hi[a] + pad, ctr[c][a], half[c][a]);
    float4* N = &out.nodes[4 * (size_t)self];
    N[0] = float4{ctr[0][0], half[0][0], ctr[0][1], half[0][1]};
    N[1] = float4{ctr[1][0], half[1][0], ctr[1][1], half[1][1]};
    N[2] = float4{ctr[0][2], half[0][2], ctr[1][2], half[1][2]};
    float r0, r1;
    std::memcpy(&r0, &refs[0], 4);
    std::memcpy(&r1, &refs[1], 4);
    N[3] = float4{r0, r1, 0.f, 0.f};
    emit_qnode(self, *cb[0], *cb[1], refs[0], refs[1]);
    return self;
  };
  if (B.surfaces.empty()) {
    // no surfaces: the kernels skip traversal when n_surface_prims == 0; keep one inert node
    out.nodes = {float4{0.f, 0.f, 0.f, 0.f}, float4{0.f, 0.f, 0.f, 0.f}, float4{0.f, 0.f, 0.f, 0.f}, float4{0.f, 0.f, 0.f, 0.f}};
    out.qnodes = {uint4{0u, 0u, 0u, 0u}, uint4{0u, 0u, 0u, 0u}};
  } else if (bvh.nodes[0].left < 0) {
    // a single leaf: both children of the root reference it (the second test is a no-op by the
    // tie rule -- the reference's BvhNode::new duplicates a lone object the same way, hittable.rs:161-162)
    const Box& bx = bvh.nodes[0].box;
    float c3[3], h3[3];
    for (int a = 0; a < 3; a++) box_center_half(bx.lo[a] - pad, bx.hi[a] + pad, c3[a], h3[a]);
    const float4 n0 = float4{c3[0], h3[0], c3[1], h3[1]};
    const float lz = c3[2], hz = h3[2];
    const int ref0 = leaf_ref(0, bvh.nodes[0].count);
    float r0;
    std::memcpy(&r0, &ref0, 4);
    out.nodes = {n0, n0, float4{lz, hz, lz, hz}, float4{r0, r0, 0.f, 0.f}};
    emit_qnode(0, bx, bx, ref0, ref0);
  } else {
    emit_node(0);
  }

  // Scenes whose primitives are tiny against the scene extent (book-1: 0.2-radius spheres on a
  // 1000-radius ground) lose too much to the grid: those keep the 64-byte fp32 nodes.
  // MEASURED on c4 (B200): 35.1 ms extend per step with the 32-byte nodes vs 34.2 ms with the fp32 nodes --
  // the L1 wavefronts fall as intended, but the loop is bound by instruction issue and the 12 extra PRMT
  // per visit cost more than the loads save.  Kept as an opt-in arm (RTB_FLAG_QNODES) for scenes whose
  // trees do not fit L1; the surface-area test still vetoes it where the grid is too coarse.
  out.use_qnodes = 0;
  if (d.flags & RTB_FLAG_QNODES) out.use_qnodes = (!B.surfaces.empty() && sa_quant <= 1.03 * sa_exact) ? 1 : 0;
  // ---- BVH4: collapse every other level of the emitted BVH2 (same fp32 boxes, so the cull is the same) ----
  // A 4-wide node halves the dependent node steps per ray (c4: 12.6 -> 6.9 visits).
#if !defined(RTB_SLAB_CENTER)
  if (!B.surfaces.empty() && 3 * ((bvh.max_depth + 1) / 2) + 4 <= BVH_STACK) {
    struct Child { float lo[3], hi[3]; int ref; };
    auto child_of = [&](int node2, int c) {
      const float4* N = &out.nodes[4 * (size_t)node2];
      Child ch;
      if (c == 0) { ch.lo[0] = N[0].x; ch.hi[0] = N[0].y; ch.lo[1] = N[0].z; ch.hi[1] = N[0].w; ch.lo[2] = N[2].x; ch.hi[2] = N[2].y; std::memcpy(&ch.ref, &N[3].x, 4); }
      else { ch.lo[0] = N[1].x; ch.hi[0] = N[1].y; ch.lo[1] = N[1].z; ch.hi[1] = N[1].w; ch.lo[2] = N[2].z; ch.hi[2] = N[2].w; std::memcpy(&ch.ref, &N[3].y, 4); }
      return ch;
    };
    std::function<int(int)> emit4 = [&](int node2) -> int {
      const int self = (int)out.nodes4.size() / 8;
      out.nodes4.resize(out.nodes4.size() + 8);
      Child kids[4];
      int nk = 0;
      for (int c = 0; c < 2; c++) {
        const Child ch = child_of(node2, c);
        if (ch.ref >= 0) { kids[nk++] = child_of(ch.ref, 0); kids[nk++] = child_of(ch.ref, 1); }
        else kids[nk++] = ch;
      }
      // a lone leaf under the root is referenced by both BVH2 children: keep one
      if (nk == 2 && kids[0].ref < 0 && kids[0].ref == kids[1].ref) nk = 1;
      float v[6][4];
      int refs[4];
      const float inf = std::numeric_limits<float>::infinity();
      for (int k = 0; k < 4; k++) {
        if (k < nk) {
          for (int a = 0; a < 3; a++) { v[2 * a][k] = kids[k].lo[a]; v[2 * a + 1][k] = kids[k].hi[a]; }
          refs[k] = kids[k].ref >= 0 ? emit4(kids[k].ref) : kids[k].ref;
        } else {  // empty slot: both x planes at +inf -> the slab interval is empty for every ray
          for (int j = 0; j < 6; j++) v[j][k] = 0.f;
          v[0][k] = v[1][k] = inf;
          refs[k] = 0;
        }
      }
      float4* N = &out.nodes4[8 * (size_t)self];
      for (int j = 0; j < 6; j++) N[j] = float4{v[j][0], v[j][1], v[j][2], v[j][3]};
      float r[4];
      std::memcpy(r, refs, 16);
      N[6] = float4{r[0], r[1], r[2], r[3]};
      N[7] = float4{0.f, 0.f, 0.f, 0.f};
      return self;
    };
    emit4(0);
    // MEASURED on c4 (B200, ncu in profiles/): visits per ray fall 12.3 -> 6.9, but a 128-byte node is seven
    // load instructions per visit and the L1 data pipe -- already at 75 % with the BVH2 -- saturates
    // (81 %, long-scoreboard stalls 4.5 -> 6.1 per issue): 555 vs 487 us per launch.  Opt-in (RTB_FLAG_BVH4).
    out.use_bvh4 = 0;
    if (d.flags & RTB_FLAG_BVH4) out.use_bvh4 = 1;
  }
#endif
  if (out.nodes4.empty()) out.nodes4.assign(8, float4{0.f, 0.f, 0.f, 0.f});
  // ---- media: boundary primitives after the surfaces, in DFS order ----------------------------------
  for (size_t mi = 0; mi < B.boundaries.size(); mi++) {
    DMedium m{};
    m.first_prim = (int)out.prim_info.size();
    m.n_prims = (int)B.boundaries[mi].size();
    m.material = B.medium_material[mi];
    m.neg_inv_density = -1. / B.medium_density[mi];  // src/constant_medium.rs:26
    Box bx;
    for (const Baked& b : B.boundaries[mi]) { emit_prim(d, out, b); bx.grow(b.lo, b.hi); }
    m.cls_fast = shading_class(d, m.material);
    if (m.n_prims == 1 && B.boundaries[mi][0].kind == PRIM_SPHERE && !(B.boundaries[mi][0].flags & PRIM_FLAG_MOVING)) {
      m.cls_fast |= 0x100;
      const double* sp = B.boundaries[mi][0].payload;  // cx cy cz r
      for (int a = 0; a < 3; a++) m.sphere[a] = (float)sp[a];
      // fp32 evaluation of |p - c|^2 at scene magnitude M errs by ~1e-6 (M + r)^2; 1e-3 r^2 covers it for r >= M / 20,
      // a smaller sphere loses the shortcut near its surface only (the margin below is then subtracted explicitly)
      const double r2 = sp[3] * sp[3];
      m.sphere[3] = round_down(std::max(0., r2 * (1. - 1e-3) - 4e-6 * (M + std::fabs(sp[3])) * (M + std::fabs(sp[3]))));
    }
    {
      bool all_quads = m.n_prims > 0;
      for (const Baked& b : B.boundaries[mi]) all_quads = all_quads && b.kind == PRIM_QUAD;
      if (all_quads && !(d.flags & RTB_FLAG_NO_BOX_SCAN)) m.cls_fast |= 0x200;
    }
    if ((m.cls_fast & 0x200) && m.n_prims == 6) {
      // the six quads of a make_box (src/object.rs:509-560) after baking: an oriented box?  Axes from the first quad,
      // every quad's normal along one axis, its corners on one face and spanning the other two extents.
      const std::vector<Baked>& q6 = B.boundaries[mi];
      auto vec = [](const double* p) { return D3{p[0], p[1], p[2]}; };
      const D3 u0 = vec(q6[0].payload + 7), v0 = vec(q6[0].payload + 10);
      const D3 ax[3] = {unit(u0), unit(v0), unit(cross(u0, v0))};
      const double scale = std::max(length(u0), length(v0));
      bool ok = std::fabs(dot(ax[0], ax[1])) < 1e-9;
      double lo3[3] = {kInf, kInf, kInf}, hi3[3] = {-kInf, -kInf, -kInf};
      for (const Baked& b : q6) {
        const D3 q = vec(b.payload + 4), u = vec(b.payload + 7), v = vec(b.payload + 10);
        for (const D3& c : {q, q + u, q + v, q + u + v})
          for (int k = 0; k < 3; k++) { lo3[k] = std::min(lo3[k], dot(c, ax[k])); hi3[k] = std::max(hi3[k], dot(c, ax[k])); }
      }
      int faces = 0;
      int face_of[6] = {0, 0, 0, 0, 0, 0};
      for (const Baked& b : q6) {
        const D3 n = vec(b.payload), q = vec(b.payload + 4), u = vec(b.payload + 7), v = vec(b.payload + 10);
        int k = -1;
        for (int a = 0; a < 3; a++)
          if (std::fabs(std::fabs(dot(n, ax[a])) - 1.) < 1e-9) k = a;
        if (k < 0) { ok = false; break; }
        const double tol = 1e-9 * std::max(scale, 1.);
        double cmin[3] = {kInf, kInf, kInf}, cmax[3] = {-kInf, -kInf, -kInf};
        for (const D3& c : {q, q + u, q + v, q + u + v})
          for (int a = 0; a < 3; a++) { cmin[a] = std::min(cmin[a], dot(c, ax[a])); cmax[a] = std::max(cmax[a], dot(c, ax[a])); }
        const bool on_lo = std::fabs(cmin[k] - lo3[k]) < tol && std::fabs(cmax[k] - lo3[k]) < tol;
        const bool on_hi = std::fabs(cmin[k] - hi3[k]) < tol && std::fabs(cmax[k] - hi3[k]) < tol;
        if (!(on_lo || on_hi)) { ok = false; break; }
        for (int a = 0; a < 3; a++)
          if (a != k && (std::fabs(cmin[a] - lo3[a]) > tol || std::fabs(cmax[a] - hi3[a]) > tol)) ok = false;
        face_of[&b - q6.data()] = 2 * k + (on_hi ? 1 : 0);
        faces |= 1 << (2 * k + (on_hi ? 1 : 0));
      }
      if (ok && faces == 63) {
        m.cls_fast |= 0x400;
        // the boundary records in FACE order (2 k + the +axis_k side): medium_obb names faces by this index
        {
          const size_t f0 = (size_t)m.first_prim;
          const std::vector<double> prims6(out.prims.begin() + f0 * PRIM_DOUBLES, out.prims.begin() + (f0 + 6) * PRIM_DOUBLES);
          const std::vector<DPre> pre6(out.pre.begin() + f0, out.pre.begin() + f0 + 6);
          const std::vector<int4> info6(out.prim_info.begin() + f0, out.prim_info.begin() + f0 + 6);
          for (int i = 0; i < 6; i++) {
            const size_t dst = f0 + face_of[i];
            std::copy(prims6.begin() + (size_t)i * PRIM_DOUBLES, prims6.begin() + (size_t)(i + 1) * PRIM_DOUBLES, out.prims.begin() + dst * PRIM_DOUBLES);
            out.pre[dst] = pre6[i];
            out.prim_info[dst] = info6[i];
          }
        }
        const double margin = 1e-5 * M;  // fp32 evaluation of the local coordinates at scene magnitude M errs by ~1e-6 M
        for (int k = 0; k < 3; k++) {
          const double centre_k = 0.5 * (lo3[k] + hi3[k]), half = 0.5 * (hi3[k] - lo3[k]);
          m.obb_half_out[k] = round_up(half + margin);
          m.obb_half_in[k] = round_down(std::max(0., half - margin));
          for (int a = 0; a < 3; a++) {
            const double axv[3] = {ax[k].x, ax[k].y, ax[k].z};
            m.obb_ax[k][a] = (float)axv[a];
            m.obb_c[a] += (float)0.;  // (accumulated below in f64)
          }
          (void)centre_k;
        }
        double c3[3] = {0., 0., 0.};
        for (int k = 0; k < 3; k++) {
          const double centre_k = 0.5 * (lo3[k] + hi3[k]);
          c3[0] += centre_k * ax[k].x; c3[1] += centre_k * ax[k].y; c3[2] += centre_k * ax[k].z;
        }
        for (int a = 0; a < 3; a++) m.obb_c[a] = (float)c3[a];
      }
    }
    for (int a = 0; a < 3; a++) { m.lo[a] = round_down(bx.lo[a] - pad); m.hi[a] = round_up(bx.hi[a] + pad); }
    {
      const double ex = (double)m.hi[0] - m.lo[0], ey = (double)m.hi[1] - m.lo[1], ez = (double)m.hi[2] - m.lo[2];
      m.diag = round_up(std::sqrt(ex * ex + ey * ey + ez * ez) * (1. + 1e-6));
    }
    out.media.push_back(m);
  }
  // ---- features in use (shade kernel specialisation) ----------------------------------------------------
  out.spec_bits = 0;
  if (!out.media.empty()) out.spec_bits |= SPEC_MEDIA;
  for (const DMedium& m : out.media) {
    if (m.cls_fast & 0x200) out.spec_bits |= SPEC_BOXSCAN;
    if (!(m.cls_fast & 0x100)) out.spec_bits |= SPEC_GENERIC_MEDIA;
  }
  if (!out.lights.empty()) out.spec_bits |= SPEC_LIGHTS;
  if (out.multi_leaf) out.spec_bits |= SPEC_MULTI_LEAF;
  for (const DTexture& t : out.textures)
    if (t.kind != TEX_SOLID) out.spec_bits |= SPEC_TEXTURES;
  // deferred shading of the textured classes (wavefront.cu, k_wf_shade_rare) needs every non-solid texture to hang
  // off a Lambertian SURFACE material: then "class LAMBERT_TEX / NOISE" and "evaluates a texture" are the same set
  out.defer_ok = 1;
  auto root_solid = [&](int mat) {
    const DMaterial& m = out.materials[mat];
    const bool textured = m.kind == MAT_LAMBERTIAN || m.kind == MAT_DIFFUSE_LIGHT || m.kind == MAT_ISOTROPIC;
    return !textured || out.textures[m.texture].kind == TEX_SOLID;
  };
  for (size_t i = 0; i < out.materials.size(); i++)
    if (!root_solid((int)i) && out.materials[i].kind != MAT_LAMBERTIAN) out.defer_ok = 0;
  for (const DMedium& m : out.media)
    if (!root_solid(m.material)) out.defer_ok = 0;
  for (int i = 0; i < out.n_surface_prims; i++) {
    const int4& info = out.prim_info[i];
    if (info.y >= 0 && info.y < (int)out.materials.size() && out.materials[info.y].needs_uv)
      out.spec_bits |= (info.x & 0xFF) == PRIM_QUAD ? SPEC_QUAD_UV : SPEC_SPHERE_UV;
  }
  return RTB_OK;
}

}  // namespace rtb
