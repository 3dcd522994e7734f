// tests/emu/emu.cpp -- DEVELOPMENT AID, NOT PRODUCT and NOT a fallback.
// Compiles the device header (csrc/rtb_device.cuh) and the host flattener as plain C++ so that the
// kernel logic can be exercised in a container without a GPU (tests/test_emu_*.py compare it with
// the oracle).  librtb200.so never links or calls this; it exists only to catch logic bugs before
// GPU time is spent.  Moving-sphere time is fp32 here exactly as in the render kernels.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../surely_raytracing_b200/csrc/flatten.h"
#include "../../surely_raytracing_b200/csrc/rtb_device.cuh"

using namespace rtb;

namespace {
struct Emu {
  HostScene host;
  DScene dev{};
  std::vector<double2> prims2;
};
thread_local std::string g_err;
}  // namespace

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

void* emu_scene_create(const RtbSceneDesc* d) {
  Emu* e = new Emu();
  std::string err;
  if (flatten_scene(*d, e->host, err) != RTB_OK) { g_err = err; delete e; return nullptr; }
  const HostScene& h = e->host;
  e->prims2.resize(h.prims.size() / 2);
  std::memcpy(e->prims2.data(), h.prims.data(), h.prims.size() * sizeof(double));
  DScene& D = e->dev;
  D.nodes = h.nodes.data(); D.prims = e->prims2.data(); D.prim_info = h.prim_info.data(); D.xforms = h.xforms.data();
  D.media = h.media.data(); D.materials = h.materials.data(); D.textures = h.textures.data(); D.texels = h.texels.data();
  D.perlin_vec = h.perlin_vec.data(); D.perlin_perm = h.perlin_perm.data(); D.lights = h.lights.data();
  D.n_nodes = (int)h.nodes.size() / 4; D.n_surface_prims = h.n_surface_prims; D.n_prims = (int)h.prim_info.size();
  D.n_media = (int)h.media.size(); D.n_lights = (int)h.lights.size(); D.bvh_depth = h.bvh_depth;
  D.n_materials = (int)h.materials.size(); D.n_textures = (int)h.textures.size();
  D.qnodes = h.qnodes.data(); D.use_qnodes = h.use_qnodes;
  D.nodes4 = h.nodes4.data(); D.use_bvh4 = h.use_bvh4;
  D.spec_bits = h.spec_bits;
  D.pre = h.pre.data(); D.scene_mag = h.scene_mag;
  D.multi_leaf = h.multi_leaf;
  D.defer_ok = h.defer_ok;
  for (int a = 0; a < 3; a++) { D.grid_base[a] = h.grid_base[a]; D.grid_inv_cell[a] = h.grid_inv_cell[a]; D.grid_cell[a] = h.grid_cell[a]; }
  D.flags = h.flags; D.seed_lo = (uint32_t)h.seed; D.seed_hi = (uint32_t)(h.seed >> 32);
  D.cam = h.cam;
  return e;
}
void emu_scene_destroy(void* p) { delete static_cast<Emu*>(p); }

int emu_scene_info(void* p, RtbSceneInfo* info) {
  const Emu* e = static_cast<Emu*>(p);
  std::memset(info, 0, sizeof(*info));
  info->image_width = e->host.cam.width; info->image_height = e->host.cam.height;
  info->spp_used = e->host.cam.spp; info->sqrt_spp = e->host.cam.sqrt_spp; info->max_depth = e->host.cam.max_depth;
  info->n_surface_prims = e->host.n_surface_prims;
  info->n_boundary_prims = (int)e->host.prim_info.size() - e->host.n_surface_prims;
  info->n_media = (int)e->host.media.size(); info->n_bvh_nodes = (int)e->host.nodes.size() / 4;
  info->n_lights = (int)e->host.lights.size(); info->bvh_depth = e->host.bvh_depth; info->device = -1;
  return 0;
}

// the body of k_render_mega for every pixel, sequentially
int emu_render(void* p, long long s_begin, long long s_end, double* rgb, unsigned long long* stats6) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  DStats st = {};
  const int n = S.cam.width * S.cam.height;
  for (int pixel = 0; pixel < n; pixel++) {
    double sr = 0, sg = 0, sb = 0;
    for (long long s = s_begin; s < s_end; s++) {
      PathState ps;
      generate_primary(S, (uint32_t)pixel, (uint32_t)s, ps);
      st.paths++;
      float Lr = 0, Lg = 0, Lb = 0;
      bool alive = true;
      while (alive) {
        Event ev;
        st.segments++;
        extend<true>(S, ps, ev, &st);
        alive = shade(S, ps, ev, Lr, Lg, Lb, &st, true);
      }
      const bool finite = (fabsf(Lr) < 3.0e38f) && (fabsf(Lg) < 3.0e38f) && (fabsf(Lb) < 3.0e38f);
      if (finite || (S.flags & 2u)) { sr += Lr; sg += Lg; sb += Lb; } else st.nonfinite++;
    }
    rgb[3 * pixel] += (double)(float)sr; rgb[3 * pixel + 1] += (double)(float)sg; rgb[3 * pixel + 2] += (double)(float)sb;
  }
  if (stats6) { stats6[0] = st.paths; stats6[1] = st.segments; stats6[2] = st.node_visits; stats6[3] = st.prim_tests; stats6[4] = st.medium_probes; stats6[5] = st.nonfinite; }
  return 0;
}

int emu_trace(void* p, const RtbRay* rays, long long n, unsigned flags, RtbHit* hits) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  for (long long i = 0; i < n; i++) {
    Ray r;
    r.ox = rays[i].origin[0]; r.oy = rays[i].origin[1]; r.oz = rays[i].origin[2];
    r.dx = rays[i].direction[0]; r.dy = rays[i].direction[1]; r.dz = rays[i].direction[2];
    r.time = rays[i].time;
    Hit best;
    hit_reset(best);
    if (flags & RTB_TRACE_BRUTE_FORCE) closest_surface_brute(S, r, rays[i].t_min, best);
    else if (S.n_surface_prims > 0) closest_surface<false>(S, r, rays[i].t_min, best, nullptr);
    complete_hit(S, r, best, hits[i]);
  }
  return 0;
}

// the candidate scheme of the wavefront pipeline (closest_candidates -> exact resolution of the survivors, exact
// re-trace on overflow), ray by ray.  Directions are taken as given (f64, like the queue's primary records);
// counts[0] = overflows, counts[1] = candidates resolved, counts[2] = rays with two candidates.
int emu_trace_candidates(void* p, const RtbRay* rays, long long n, RtbHit* hits, unsigned long long* counts3) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  unsigned long long overflow = 0, resolved = 0, two = 0;
  for (long long i = 0; i < n; i++) {
    if (rays[i].t_min != 0.0001) { g_err = "the wavefront harness traces radiance rays: t_min must be 1e-4"; return -1; }
    Ray r;
    r.ox = rays[i].origin[0]; r.oy = rays[i].origin[1]; r.oz = rays[i].origin[2];
    r.dx = rays[i].direction[0]; r.dy = rays[i].direction[1]; r.dz = rays[i].direction[2];
    r.time = (double)(float)rays[i].time;
    Hit best;
    hit_reset(best);
    if (S.n_surface_prims > 0) {
      Cands C;
      const float mag = fmaxf(S.scene_mag, (float)(fabs(r.ox) + fabs(r.oy) + fabs(r.oz)));
      if (closest_candidates<false>(S, r, mag, C, nullptr)) {
        int r0, r1;
        cands_record(C, r0, r1);
        resolve_candidates<true>(S, r0, r1, r, 0.0001, best);
        resolved += (C.c0 < 0) + (C.c1 < 0);
        two += C.c1 < 0;
      } else {
        overflow++;
        closest_surface<false>(S, r, 0.0001, best, nullptr);
      }
    }
    complete_hit(S, r, best, hits[i]);
  }
  if (counts3) { counts3[0] = overflow; counts3[1] = resolved; counts3[2] = two; }
  return 0;
}

int emu_medium_interval(void* p, int medium, const RtbRay* rays, long long n, double* t0, double* t1) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  for (long long i = 0; i < n; i++) {
    Ray r;
    r.ox = rays[i].origin[0]; r.oy = rays[i].origin[1]; r.oz = rays[i].origin[2];
    r.dx = rays[i].direction[0]; r.dy = rays[i].direction[1]; r.dz = rays[i].direction[2];
    r.time = rays[i].time;
    double a, b;
    if (medium_interval(S, S.media[medium], r, a, b)) { t0[i] = a; t1[i] = b; } else { t0[i] = t1[i] = NAN; }
  }
  return 0;
}

int emu_eval_texture(void* p, int texture, const double* uvp, long long n, double* rgb) {
  const DScene& S = static_cast<Emu*>(p)->dev;
  for (long long i = 0; i < n; i++) {
    const V3 c = texture_value(S, texture, (float)uvp[5 * i], (float)uvp[5 * i + 1], uvp[5 * i + 2], uvp[5 * i + 3], uvp[5 * i + 4]);
    rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
  }
  return 0;
}

int emu_eval_light_pdf(void* p, const double* od, long long n, double* pdf) {
  const DScene& S = static_cast<Emu*>(p)->dev;
  for (long long i = 0; i < n; i++) {
    Ray probe;
    probe.ox = od[6 * i]; probe.oy = od[6 * i + 1]; probe.oz = od[6 * i + 2];
    probe.dx = od[6 * i + 3]; probe.dy = od[6 * i + 4]; probe.dz = od[6 * i + 5];
    probe.time = 0.;
    double sum = 0.;
    for (int k = 0; k < S.n_lights; k++) sum += light_pdf_one(S.lights[k], probe);
    pdf[i] = sum * (1. / (double)S.n_lights);
  }
  return 0;
}

}  // extern "C"

// diagnostic: per-segment node-visit counts over a render pass; records the rays whose traversal
// visited more than `threshold` nodes (first `cap` of them) and returns how many there were.
extern "C" long long emu_slow_rays(void* p, long long s_begin, long long s_end, int pixel_stride, unsigned long long threshold,
                                   RtbRay* out, unsigned long long* out_visits, int cap, unsigned long long* histogram16) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  const int n = S.cam.width * S.cam.height;
  long long found = 0;
  for (int pixel = 0; pixel < n; pixel += pixel_stride)
    for (long long s = s_begin; s < s_end; s++) {
      PathState ps;
      generate_primary(S, (uint32_t)pixel, (uint32_t)s, ps);
      float Lr = 0, Lg = 0, Lb = 0;
      bool alive = true;
      while (alive) {
        DStats st = {};
        Event ev;
        const Ray r = ps.ray;
        extend<true>(S, ps, ev, &st);
        int b = 0;
        while ((1ull << b) < st.node_visits && b < 15) b++;
        histogram16[b]++;
        if (st.node_visits > threshold) {
          if (found < cap) {
            RtbRay& o = out[found];
            o.origin[0] = r.ox; o.origin[1] = r.oy; o.origin[2] = r.oz;
            o.direction[0] = r.dx; o.direction[1] = r.dy; o.direction[2] = r.dz;
            o.time = r.time; o.t_min = 1e-4;
            out_visits[found] = st.node_visits;
          }
          found++;
        }
        DStats dummy = {};
        alive = shade(S, ps, ev, Lr, Lg, Lb, &dummy, false);
      }
    }
  return found;
}


// features the flattener found in the scene (device_scene.h SPEC_*), and the leaf-reference codec
extern "C" int emu_spec_bits(void* p) { return static_cast<Emu*>(p)->host.spec_bits; }
extern "C" int emu_defer_ok(void* p) { return static_cast<Emu*>(p)->host.defer_ok; }
extern "C" int emu_leaf_roundtrip(int first, int count, int kind_bits) {
  const int ref = leaf_make(first, count, kind_bits);
  return ref < 0 && leaf_first(ref) == first && leaf_count(ref) == count && leaf_kind_bits(ref) == kind_bits;
}
// cls_fast of medium `mi` (bit 8: sphere boundary, bit 9: quads only, bit 10: oriented box recognised)
extern "C" int emu_medium_flags(void* p, int mi) {
  const HostScene& h = static_cast<Emu*>(p)->host;
  return mi >= 0 && mi < (int)h.media.size() ? h.media[mi].cls_fast : -1;
}
// child references of inner node `node` of the BVH2 (device numbering)
extern "C" void emu_node_children(void* p, int node, int* out2) {
  const HostScene& h = static_cast<Emu*>(p)->host;
  std::memcpy(out2, &h.nodes[4 * (size_t)node + 3], 8);
}
// every leaf of the BVH2 carries the kind bits of its (single) primitive; returns the number of violations
extern "C" int emu_check_leaf_refs(void* p) {
  const Emu* e = static_cast<Emu*>(p);
  const HostScene& h = e->host;
  int bad = 0;
  if (h.n_surface_prims == 0) return 0;
  for (size_t i = 0; i < h.nodes.size() / 4; i++) {
    int refs[2];
    std::memcpy(refs, &h.nodes[4 * i + 3], 8);
    for (int c = 0; c < 2; c++) {
      if (refs[c] >= 0) { bad += refs[c] >= (int)(h.nodes.size() / 4); continue; }
      const int first = leaf_first(refs[c]), count = leaf_count(refs[c]);
      if (first < 0 || first + count > h.n_surface_prims) { bad++; continue; }
      const int4 info = h.prim_info[first];
      if (leaf_kind_bits(refs[c]) == LEAF_KIND_BOX) {
        // an axis-aligned make_box: six static quads in face order, face (k, side) on plane lo[k] / hi[k] of the bounds
        // that sit in the DPre slot of the first face
        DBoxBounds bb;
        std::memcpy(&bb, &h.pre[first], sizeof(bb));
        bad += count != 6;
        for (int f = 0; f < 6 && count == 6; f++) {
          const int4 fi = h.prim_info[first + f];
          const double* P = &h.prims[(size_t)(first + f) * PRIM_DOUBLES];
          const int k = f >> 1;
          bad += (fi.x & 0xFF) != PRIM_QUAD || (fi.x & PRIM_FLAG_MOVING);
          bad += P[4 + k] != ((f & 1) ? bb.hi[k] : bb.lo[k]) || P[7 + k] != 0. || P[10 + k] != 0.;
        }
        continue;
      }
      const int want = ((info.x & 0xFF) == PRIM_QUAD ? LEAF_KIND_QUAD : 0) | ((info.x & PRIM_FLAG_MOVING) ? LEAF_KIND_MOVING : 0);
      bad += leaf_kind_bits(refs[c]) != want;
      const int mat = (info.x >> PRIM_MAT_SHIFT) & 0xFFF;
      bad += mat != 0 && mat - 1 != info.y;
    }
  }
  return bad;
}

// ---- alternative tree forms of the wavefront traversal (DScene::qnodes, DScene::nodes4): they must find
//      exactly the closest hits of the fp32 BVH2 and of a brute-force scan.  Rays as stored in the queues.
struct QRay { double ox, oy, oz; float dx, dy, dz, time; };

// conservativeness check of the 32-byte nodes: closest hit through qnodes vs the fp32 nodes vs brute force
extern "C" int emu_check_qnodes(void* p, const QRay* rays, long long n, long long brute_every, double* out6) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  const float tmin32 = __double2float_rd(0.0001);
  double visits_q = 0, visits_f = 0, mism_f = 0, mism_b = 0, n_brute = 0, unculled = 0;
  for (long long i = 0; i < n; i++) {
    Ray r;
    r.ox = rays[i].ox; r.oy = rays[i].oy; r.oz = rays[i].oz;
    r.dx = rays[i].dx; r.dy = rays[i].dy; r.dz = rays[i].dz; r.time = rays[i].time;
    Hit hq;
    hit_reset(hq);
    {
      const SlabRay sr = slab_ray_q(S, r.ox, r.oy, r.oz, rays[i].dx, rays[i].dy, rays[i].dz);
      if (sr.idx != sr.idx) unculled++;
      float tbest32 = __double2float_ru(hq.t);
      int stack[BVH_STACK], sp = 0, node = 0;
      for (;;) {
        if (node >= 0) {
          visits_q++;
          const uint4* N = S.qnodes + 2 * (size_t)node;
          float tn0, tn1;
          bool h0, h1;
          slab_box_q(N[0].x, N[0].y, N[0].z, sr, tmin32, tbest32, tn0, h0);
          slab_box_q(N[1].x, N[1].y, N[1].z, sr, tmin32, tbest32, tn1, h1);
          int ch0 = (int)N[0].w, ch1 = (int)N[1].w;
          if (h0 && h1) {
            if (tn1 < tn0) std::swap(ch0, ch1);
            stack[sp++] = ch1;
            node = ch0;
            continue;
          }
          if (h0) { node = ch0; continue; }
          if (h1) { node = ch1; continue; }
        } else {
          const int leaf = ~node;
          (void)leaf;
          test_leaf(S, node, r, 0.0001, hq);
          tbest32 = __double2float_ru(hq.t);
        }
        if (sp == 0) break;
        node = stack[--sp];
      }
    }
    Hit hf;
    hit_reset(hf);
    DStats st = {};
    closest_surface<true>(S, r, 0.0001, hf, &st);
    visits_f += (double)st.node_visits;
    if (hf.prim != hq.prim || hf.t != hq.t) mism_f++;
    if (brute_every > 0 && (i % brute_every) == 0) {
      Hit hb;
      hit_reset(hb);
      closest_surface_brute(S, r, 0.0001, hb);
      n_brute++;
      if (hb.prim != hq.prim || hb.t != hq.t) mism_b++;
    }
  }
  out6[0] = visits_q / n; out6[1] = visits_f / n; out6[2] = mism_f; out6[3] = mism_b; out6[4] = n_brute; out6[5] = unculled;
  return 0;
}


// the collapsed tree (DScene::nodes4) must find the same closest hits as the BVH2 and as brute force
extern "C" int emu_check_nodes4(void* p, const QRay* rays, long long n, long long brute_every, double* out6) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  const float tmin32 = __double2float_rd(0.0001);
  double visits4 = 0, visits2 = 0, mism2 = 0, mismb = 0, nb = 0, maxsp = 0;
  for (long long i = 0; i < n; i++) {
    Ray r;
    r.ox = rays[i].ox; r.oy = rays[i].oy; r.oz = rays[i].oz;
    r.dx = rays[i].dx; r.dy = rays[i].dy; r.dz = rays[i].dz; r.time = rays[i].time;
    Hit h4;
    hit_reset(h4);
    {
      const SlabRay sr = slab_ray(r.ox, r.oy, r.oz, rays[i].dx, rays[i].dy, rays[i].dz);
      float tbest32 = __double2float_ru(h4.t);
      int stack[256], sp = 0, node = 0;
      for (;;) {
        if (node >= 0) {
          visits4++;
          const float4* N = S.nodes4 + 8 * (size_t)node;
          const float* L[6] = {&N[0].x, &N[1].x, &N[2].x, &N[3].x, &N[4].x, &N[5].x};
          int refs[4];
          std::memcpy(refs, &N[6], 16);
          float t[4];
          bool h[4];
          for (int c = 0; c < 4; c++) slab_box(L[0][c], L[1][c], L[2][c], L[3][c], L[4][c], L[5][c], sr, tmin32, tbest32, t[c], h[c]);
          int m = -1;
          for (int c = 0; c < 4; c++) if (h[c] && (m < 0 || t[c] < t[m])) m = c;
          if (m >= 0) {
            for (int c = 0; c < 4; c++) if (h[c] && c != m) stack[sp++] = refs[c];
            if (sp > maxsp) maxsp = sp;
            node = refs[m];
            continue;
          }
        } else {
          const int leaf = ~node;
          (void)leaf;
          test_leaf(S, node, r, 0.0001, h4);
          tbest32 = __double2float_ru(h4.t);
        }
        if (sp == 0) break;
        node = stack[--sp];
      }
    }
    Hit h2;
    hit_reset(h2);
    DStats st = {};
    closest_surface<true>(S, r, 0.0001, h2, &st);
    visits2 += (double)st.node_visits;
    if (h2.prim != h4.prim || h2.t != h4.t) mism2++;
    if (brute_every > 0 && (i % brute_every) == 0) {
      Hit hb;
      hit_reset(hb);
      closest_surface_brute(S, r, 0.0001, hb);
      nb++;
      if (hb.prim != h4.prim || hb.t != h4.t) mismb++;
    }
  }
  out6[0] = visits4 / n; out6[1] = visits2 / n; out6[2] = mism2; out6[3] = mismb; out6[4] = nb; out6[5] = maxsp;
  return 0;
}
