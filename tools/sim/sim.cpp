// tools/sim/sim.cpp -- DEVELOPMENT AID (not product, not a fallback): a CPU model of how the lanes of
// a warp are occupied in k_wf_extend, used to rank scheduling policies before GPU time is spent.
//
//  1. sim_collect: emulates the wavefront pipeline (generate -> extend -> block-sorted shade -> dense
//     append) on the host build of the device header and records the ray queue of chosen iterations,
//     in queue order, so the simulated warps see realistic mixtures of bounce depths.
//  2. sim_extend: replays k_wf_extend's control flow for W round-robin persistent warps over such a
//     queue under a policy (fetch threshold, postponed-leaf slots, stale-pop culling, ...), counting
//     warp-level loop trips, active lanes and a slot-cost estimate.
#include <algorithm>
#include <cstdio>

#include "../../tests/emu/emu.cpp"

namespace {

int shading_class_of(const DScene& S, const Event& ev) {
  if (!(ev.t < RTB_INF)) return CLS_MISS;
  if (ev.medium >= 0) return S.media[ev.medium].cls_fast & 0xF;
  return (S.prim_info[ev.prim].x >> PRIM_CLASS_SHIFT) & 0xF;
}

struct Lane {
  bool have = false;
  Ray r;
  SlabRay sr;
  float tbest32 = 0.f;
  Hit best;
  int stack[BVH_STACK];
  float stack_t[BVH_STACK];
  int sp = 0, node = 0x7FFFFFFF;
  int leaf[4] = {0, 0, 0, 0};
  int n_leaf = 0;
};

constexpr int DONE = 0x7FFFFFFF;

struct Policy {
  int threshold;      // refill when fewer lanes hold a ray
  int leaf_slots;     // postponed leaves per lane (1 = shipped kernel)
  int stale_cull;     // 1: stack entries carry their entry distance and are dropped when beyond tbest
  int break_mode;     // 0: leave the inner loop when every looping lane holds >= 1 leaf (shipped)
                      // 1: ... when every looping lane has all slots full
                      // 2: ... when at least `break_count` lanes of the warp hold a leaf or are idle
  int break_count;
  int n_warps;
  int cost_inner, cost_quad, cost_sphere, cost_fetch, cost_outer;
};

struct SimOut {
  double rays, inner_trips, inner_lane_sum, leaf_rounds, leaf_lane_sum, node_visits, prim_tests, fetches, outer_trips, cost;
  double stale_skipped;
  double distinct_nodes;  // sum over inner trips of the number of distinct nodes the active lanes visit
};

}  // namespace

extern "C" {

// Emulate the pipeline for strata [s_begin, s_end) with `cap` path slots; rays of iterations
// [it_lo, it_hi) are appended to `out` (at most out_cap), returns the number written; iteration
// boundaries go to it_offsets (it_hi - it_lo + 1 entries).
long long sim_collect(void* p, int cap, long long s_begin, long long s_end, int it_lo, int it_hi, QRay* out,
                      long long out_cap, long long* it_offsets) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  const uint32_t tiles_x = (uint32_t)(S.cam.width + 7) >> 3, tiles_y = (uint32_t)(S.cam.height + 3) >> 2;
  const unsigned long long padded = (unsigned long long)tiles_x * tiles_y * 32ull;
  const unsigned long long total = padded * (unsigned long long)(s_end - s_begin);
  unsigned long long next_path = 0;
  std::vector<PathState> in, outq;
  std::vector<char> pad_in, pad_out;
  long long n_written = 0;
  for (int iter = 0;; iter++) {
    // generate: top up
    while ((int)outq.size() < cap && next_path < total) {
      const unsigned long long pid = next_path++;
      const uint32_t sample = (uint32_t)(s_begin + (long long)(pid / padded));
      const uint32_t idx = (uint32_t)(pid % padded), tile = idx >> 5, lane = idx & 31u;
      const uint32_t x = (tile % tiles_x) * 8u + (lane & 7u), y = (tile / tiles_x) * 4u + (lane >> 3);
      if (x < (uint32_t)S.cam.width && y < (uint32_t)S.cam.height) {
        PathState ps;
        generate_primary(S, y * (uint32_t)S.cam.width + x, sample, ps);
        ps.ray.dx = (double)(float)ps.ray.dx; ps.ray.dy = (double)(float)ps.ray.dy; ps.ray.dz = (double)(float)ps.ray.dz;
        ps.ray.time = (double)(float)ps.ray.time;
        outq.push_back(ps);
      }
    }
    in.swap(outq);
    outq.clear();
    if (in.empty()) break;
    if (iter >= it_lo && iter < it_hi) {
      it_offsets[iter - it_lo] = n_written;
      for (const PathState& ps : in) {
        if (n_written >= out_cap) break;
        QRay q;
        q.ox = ps.ray.ox; q.oy = ps.ray.oy; q.oz = ps.ray.oz;
        q.dx = (float)ps.ray.dx; q.dy = (float)ps.ray.dy; q.dz = (float)ps.ray.dz; q.time = (float)ps.ray.time;
        out[n_written++] = q;
      }
      it_offsets[iter - it_lo + 1] = n_written;
    }
    if (iter + 1 >= it_hi) break;
    // extend + block-sorted shade
    std::vector<Event> evs(in.size());
    DStats st = {};
    for (size_t i = 0; i < in.size(); i++) extend<false>(S, in[i], evs[i], &st);
    for (size_t b0 = 0; b0 < in.size(); b0 += 256) {
      const size_t b1 = std::min(in.size(), b0 + 256);
      std::vector<int> order;
      for (int k = 0; k < NUM_CLASSES; k++)
        for (size_t i = b0; i < b1; i++)
          if (shading_class_of(S, evs[i]) == k) order.push_back((int)i);
      for (int i : order) {
        float Lr = 0, Lg = 0, Lb = 0;
        PathState ps = in[i];
        Event ev = evs[i];
        if (shade(S, ps, ev, Lr, Lg, Lb, &st, false)) {
          ps.ray.dx = (double)(float)ps.ray.dx; ps.ray.dy = (double)(float)ps.ray.dy; ps.ray.dz = (double)(float)ps.ray.dz;
          outq.push_back(ps);
        }
      }
    }
  }
  return n_written;
}

int sim_extend(void* p, const QRay* rays, long long n, const Policy* pol_, SimOut* o) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  const Policy pol = *pol_;
  *o = SimOut{};
  const float tmin32 = __double2float_rd(0.0001);
  long long cursor = 0;
  std::vector<std::vector<Lane>> warps(pol.n_warps, std::vector<Lane>(32));
  std::vector<char> warp_done(pol.n_warps, 0);
  int n_done = 0;
  auto pop = [&](Lane& L) {
    for (;;) {
      if (L.sp == 0) { L.node = DONE; return; }
      L.sp--;
      if (pol.stale_cull && L.stack_t[L.sp] > L.tbest32) { o->stale_skipped++; continue; }
      L.node = L.stack[L.sp];
      return;
    }
  };
  while (n_done < pol.n_warps) {
    for (int w = 0; w < pol.n_warps; w++) {
      if (warp_done[w]) continue;
      std::vector<Lane>& W = warps[w];
      o->outer_trips++;
      o->cost += pol.cost_outer;
      // ---- fetch
      int have = 0;
      for (auto& L : W) have += L.have;
      if (have < pol.threshold && cursor < n) {
        bool any = false;
        for (auto& L : W) {
          if (L.have || cursor >= n) continue;
          const QRay& q = rays[cursor++];
          L.r.ox = q.ox; L.r.oy = q.oy; L.r.oz = q.oz;
          L.r.dx = q.dx; L.r.dy = q.dy; L.r.dz = q.dz; L.r.time = q.time;
          L.sr = slab_ray(q.ox, q.oy, q.oz, q.dx, q.dy, q.dz);
          hit_reset(L.best);
          L.tbest32 = __double2float_ru(L.best.t);
          L.sp = 0; L.n_leaf = 0; L.node = S.n_surface_prims > 0 ? 0 : DONE;
          L.have = true;
          any = true;
          o->rays++;
        }
        if (any) { o->fetches++; o->cost += pol.cost_fetch; }
      }
      have = 0;
      for (auto& L : W) have += L.have;
      if (have == 0) { warp_done[w] = 1; n_done++; continue; }
      // ---- inner loop
      int entered = 0;
      for (auto& L : W) entered += (L.have && L.node >= 0 && L.node != DONE);
      for (;;) {
        int active = 0;
        for (auto& L : W) active += (L.have && L.node >= 0 && L.node != DONE);
        if (active == 0) break;
        o->inner_trips++; o->inner_lane_sum += active; o->cost += pol.cost_inner;
        {
          int seen[32], ns = 0;
          for (auto& L : W) {
            if (!(L.have && L.node >= 0 && L.node != DONE)) continue;
            bool dup = false;
            for (int k = 0; k < ns; k++) dup |= (seen[k] == L.node);
            if (!dup) seen[ns++] = L.node;
          }
          o->distinct_nodes += ns;
        }
        for (auto& L : W) {
          if (!(L.have && L.node >= 0 && L.node != DONE)) continue;
          o->node_visits++;
          const float4* N = S.nodes + 4 * (size_t)L.node;
          const float4 n0 = N[0], n1 = N[1], n2 = N[2], n3 = N[3];
          float tn0, tn1;
          bool h0, h1;
          slab_box(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, L.sr, tmin32, L.tbest32, tn0, h0);
          slab_box(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, L.sr, tmin32, L.tbest32, tn1, h1);
          int ch0 = __float_as_int(n3.x), ch1 = __float_as_int(n3.y);
          if (h0 && h1) {
            if (tn1 < tn0) { std::swap(ch0, ch1); std::swap(tn0, tn1); }
            L.stack_t[L.sp] = tn1;
            L.stack[L.sp++] = ch1;
            L.node = ch0;
          } else if (h0) L.node = ch0;
          else if (h1) L.node = ch1;
          else pop(L);
          while (L.node < 0 && L.n_leaf < pol.leaf_slots) {  // postpone and continue
            L.leaf[L.n_leaf++] = L.node;
            pop(L);
          }
        }
        // break condition
        int looping = 0, looping_free = 0, looping_empty = 0, ready = 0;
        for (auto& L : W) {
          const bool loop = L.have && L.node >= 0 && L.node != DONE;
          looping += loop;
          looping_free += loop && L.n_leaf < pol.leaf_slots;
          looping_empty += loop && L.n_leaf == 0;
          ready += (!loop && L.have && (L.n_leaf > 0 || L.node < 0));
        }
        if (pol.break_mode == 0 && looping_empty == 0) break;
        if (pol.break_mode == 1 && looping_free == 0) break;
        if (pol.break_mode == 2 && (looping_empty == 0 || ready >= pol.break_count)) break;
        if (pol.break_mode == 3 && (looping_empty == 0 || entered - looping >= pol.break_count)) break;
      }
      // ---- leaf phase: postponed slots in order, then the current node if it is a leaf
      for (;;) {
        int lanes = 0, nq = 0, ns = 0;
        for (auto& L : W) {
          if (!L.have) continue;
          if (L.n_leaf == 0 && L.node < 0) { L.leaf[L.n_leaf++] = L.node; pop(L); }  // node leaf moves into a slot
          if (L.n_leaf == 0) continue;
          lanes++;
          const int first = leaf_first(L.leaf[0]), count = leaf_count(L.leaf[0]);
          for (int i = 0; i < count; i++) {
            o->prim_tests++;
            ((S.prim_info[first + i].x & 0xFF) == PRIM_QUAD ? nq : ns)++;
          }
          test_leaf(S, L.leaf[0], L.r, 0.0001, L.best);
          for (int k = 1; k < L.n_leaf; k++) L.leaf[k - 1] = L.leaf[k];
          L.n_leaf--;
          L.tbest32 = __double2float_ru(L.best.t);
        }
        if (lanes == 0) break;
        o->leaf_rounds++; o->leaf_lane_sum += lanes;
        o->cost += (nq ? pol.cost_quad : 0) + (ns ? pol.cost_sphere : 0);
      }
      for (auto& L : W)
        if (L.have && L.node == DONE && L.n_leaf == 0) L.have = false;
    }
  }
  return 0;
}

}  // extern "C"

// ---- BVH4 model: collapse every other level of the BVH2, replay the same lane logic --------------------------
namespace {
struct Node4 { float lo[4][3], hi[4][3]; int ref[4]; int n; };
struct Lane4 {
  bool have = false;
  Ray r;
  float ox, oy, oz, idx, idy, idz;
  float tbest32 = 0.f;
  Hit best;
  int stack[128];
  int sp = 0, node = DONE, leaf = 0;
};
}  // namespace

extern "C" int sim_extend4(void* p, const QRay* rays, long long n, const Policy* pol_, SimOut* o) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  const Policy pol = *pol_;
  *o = SimOut{};
  // collapse: node4 index = node2 index (only those reached are used)
  std::vector<Node4> N4(S.n_nodes);
  auto child_box = [&](int node2, int c, float* lo, float* hi, int& ref) {
    const float4* N = S.nodes + 4 * (size_t)node2;
    if (c == 0) { lo[0] = N[0].x; hi[0] = N[0].y; lo[1] = N[0].z; hi[1] = N[0].w; lo[2] = N[2].x; hi[2] = N[2].y; ref = __float_as_int(N[3].x); }
    else { lo[0] = N[1].x; hi[0] = N[1].y; lo[1] = N[1].z; hi[1] = N[1].w; lo[2] = N[2].z; hi[2] = N[2].w; ref = __float_as_int(N[3].y); }
  };
  for (int i = 0; i < S.n_nodes; i++) {
    Node4& q = N4[i];
    q.n = 0;
    for (int c = 0; c < 2; c++) {
      float lo[3], hi[3];
      int ref;
      child_box(i, c, lo, hi, ref);
      if (ref >= 0) {
        for (int g = 0; g < 2; g++) { child_box(ref, g, q.lo[q.n], q.hi[q.n], q.ref[q.n]); q.n++; }
      } else {
        for (int a = 0; a < 3; a++) { q.lo[q.n][a] = lo[a]; q.hi[q.n][a] = hi[a]; }
        q.ref[q.n++] = ref;
      }
    }
  }
  const float tmin32 = __double2float_rd(0.0001);
  long long cursor = 0;
  std::vector<std::vector<Lane4>> warps(pol.n_warps, std::vector<Lane4>(32));
  std::vector<char> warp_done(pol.n_warps, 0);
  int n_done = 0;
  auto pop = [&](Lane4& L) { L.node = L.sp > 0 ? L.stack[--L.sp] : DONE; };
  while (n_done < pol.n_warps) {
    for (int w = 0; w < pol.n_warps; w++) {
      if (warp_done[w]) continue;
      std::vector<Lane4>& W = warps[w];
      o->outer_trips++;
      int have = 0;
      for (auto& L : W) have += L.have;
      if (have < pol.threshold && cursor < n) {
        for (auto& L : W) {
          if (L.have || cursor >= n) continue;
          const QRay& q = rays[cursor++];
          L.r.ox = q.ox; L.r.oy = q.oy; L.r.oz = q.oz;
          L.r.dx = q.dx; L.r.dy = q.dy; L.r.dz = q.dz; L.r.time = q.time;
          L.ox = (float)q.ox; L.oy = (float)q.oy; L.oz = (float)q.oz;
          L.idx = safe_rcp(q.dx); L.idy = safe_rcp(q.dy); L.idz = safe_rcp(q.dz);
          hit_reset(L.best);
          L.tbest32 = __double2float_ru(L.best.t);
          L.sp = 0; L.leaf = 0; L.node = S.n_surface_prims > 0 ? 0 : DONE;
          L.have = true;
          o->rays++;
        }
        o->fetches++;
      }
      have = 0;
      for (auto& L : W) have += L.have;
      if (have == 0) { warp_done[w] = 1; n_done++; continue; }
      int entered = 0;
      for (auto& L : W) entered += (L.have && L.node >= 0 && L.node != DONE);
      for (;;) {
        int active = 0;
        for (auto& L : W) active += (L.have && L.node >= 0 && L.node != DONE);
        if (active == 0) break;
        o->inner_trips++; o->inner_lane_sum += active;
        {
          int seen[32], ns = 0;
          for (auto& L : W) {
            if (!(L.have && L.node >= 0 && L.node != DONE)) continue;
            bool dup = false;
            for (int k = 0; k < ns; k++) dup |= (seen[k] == L.node);
            if (!dup) seen[ns++] = L.node;
          }
          o->distinct_nodes += ns;
        }
        for (auto& L : W) {
          if (!(L.have && L.node >= 0 && L.node != DONE)) continue;
          o->node_visits++;
          const Node4& q = N4[L.node];
          float tn[4];
          int ref[4], nh = 0;
          for (int c = 0; c < q.n; c++) {
            const float a0 = (q.lo[c][0] - L.ox) * L.idx, a1 = (q.hi[c][0] - L.ox) * L.idx;
            const float b0 = (q.lo[c][1] - L.oy) * L.idy, b1 = (q.hi[c][1] - L.oy) * L.idy;
            const float c0 = (q.lo[c][2] - L.oz) * L.idz, c1 = (q.hi[c][2] - L.oz) * L.idz;
            const float t0 = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), tmin32));
            const float t1 = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), L.tbest32));
            if (t0 <= fmaf(fabsf(t1), 4e-6f, t1)) { tn[nh] = t0; ref[nh] = q.ref[c]; nh++; }
          }
          // sort hits near -> far (insertion), push far ones
          if (pol.stale_cull == 0) {
            for (int a = 1; a < nh; a++)
              for (int b = a; b > 0 && tn[b] < tn[b - 1]; b--) { std::swap(tn[b], tn[b - 1]); std::swap(ref[b], ref[b - 1]); }
          } else {  // (flag reused) nearest first, the rest in slot order
            int m = 0;
            for (int a = 1; a < nh; a++) if (tn[a] < tn[m]) m = a;
            if (nh > 0) { std::swap(tn[0], tn[m]); std::swap(ref[0], ref[m]); }
          }
          if (nh == 0) pop(L);
          else {
            for (int a = nh - 1; a >= 1; a--) L.stack[L.sp++] = ref[a];
            L.node = ref[0];
          }
          if (L.node < 0 && L.leaf == 0) { L.leaf = L.node; pop(L); }
        }
        int looping = 0, looping_empty = 0;
        for (auto& L : W) {
          const bool loop = L.have && L.node >= 0 && L.node != DONE;
          looping += loop;
          looping_empty += loop && L.leaf == 0;
        }
        if (looping_empty == 0) break;
        if (pol.break_mode == 3 && entered - looping >= pol.break_count) break;
      }
      for (;;) {
        int lanes = 0;
        for (auto& L : W) {
          if (!L.have) continue;
          if (L.leaf == 0 && L.node < 0) { L.leaf = L.node; pop(L); }
          if (L.leaf == 0) continue;
          lanes++;
          o->prim_tests += test_leaf(S, L.leaf, L.r, 0.0001, L.best);
          L.leaf = 0;
          L.tbest32 = __double2float_ru(L.best.t);
        }
        if (lanes == 0) break;
        o->leaf_rounds++; o->leaf_lane_sum += lanes;
      }
      for (auto& L : W)
        if (L.have && L.node == DONE && L.leaf == 0) L.have = false;
    }
  }
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Candidate scheme (rtb_device.cuh: prefilter_quad / prefilter_sphere / cands_add), checked on the host build.
//
// sim_prefilter: for every leaf test the exact BVH2 traversal of the given rays performs, classify the primitive
// with the conservative fp32 test and check the classification against the exact f64 reference-order test.
//   out: 0 tests, 1 certain miss, 2 certain hit, 3 uncertain, 4 VIOLATIONS (certain miss but exact hit, certain hit
//   but exact miss, or exact t outside [t_lo, t_hi]; an uncertain one whose exact t lies below its t_lo),
//   5 exact hits, 6 sum of relative widths of the certain-hit windows.
// ------------------------------------------------------------------------------------------------
extern "C" int sim_prefilter(void* p, const QRay* rays, long long n, double* out) {
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  const float tmin_lo = __double2float_rd(0.0001), tmin_hi = __double2float_ru(0.0001);
  for (int k = 0; k < 8; k++) out[k] = 0.;
  for (long long i = 0; i < n; i++) {
    Ray r;
    r.ox = rays[i].ox; r.oy = rays[i].oy; r.oz = rays[i].oz;
    r.dx = rays[i].dx; r.dy = rays[i].dy; r.dz = rays[i].dz; r.time = rays[i].time;
    const PfRay pr = pf_ray(r.ox, r.oy, r.oz, rays[i].dx, rays[i].dy, rays[i].dz, rays[i].time, S.scene_mag);
    Hit best;
    hit_reset(best);
    const SlabRay sr = slab_ray(r.ox, r.oy, r.oz, rays[i].dx, rays[i].dy, rays[i].dz);
    float tbest32 = __double2float_ru(best.t);
    int stack[BVH_STACK], sp = 0, node = 0;
    for (;;) {
      if (node >= 0) {
        const float4* N = S.nodes + 4 * (size_t)node;
        float tn0, tn1;
        bool h0, h1;
        slab_box(N[0].x, N[0].y, N[0].z, N[0].w, N[2].x, N[2].y, sr, tmin_lo, tbest32, tn0, h0);
        slab_box(N[1].x, N[1].y, N[1].z, N[1].w, N[2].z, N[2].w, sr, tmin_lo, tbest32, tn1, h1);
        int ch0 = __float_as_int(N[3].x), ch1 = __float_as_int(N[3].y);
        if (h0 && h1) {
          if (tn1 < tn0) std::swap(ch0, ch1);
          stack[sp++] = ch1;
          node = ch0;
          continue;
        }
        if (h0) { node = ch0; continue; }
        if (h1) { node = ch1; continue; }
      } else {
        for (int k = 0; k < leaf_count(node); k++) {
          const int pi = leaf_first(node) + k;
          const int info_x = S.prim_info[pi].x;
          const bool quad = (info_x & 0xFF) == PRIM_QUAD;
          const double2* P = S.prims + (size_t)pi * PRIM_D2;
          float t_lo = 0.f, t_hi = 0.f;
          const float inf = __int_as_float(0x7F800000);
          const int cls = quad ? prefilter_quad(P, S.pre + pi, pr, tmin_lo, tmin_hi, inf, t_lo, t_hi)
                               : prefilter_sphere(P, (info_x & PRIM_FLAG_MOVING) != 0, pr, tmin_lo, tmin_hi, inf, t_lo, t_hi);
          double t, a, b;
          const bool exact = quad ? quad_test(P, r, 0.0001, RTB_INF, t, a, b)
                                  : sphere_test(P, info_x & PRIM_FLAG_MOVING, r, r.time, 0.0001, RTB_INF, t);
          out[0]++;
          out[1 + cls]++;
          out[5] += exact;
          if (cls == PF_MISS && exact) out[4]++;
          if (cls == PF_HIT && (!exact || t < (double)t_lo || t > (double)t_hi)) out[4]++;
          if (cls == PF_UNSURE && exact && t < (double)t_lo) out[4]++;
          if (cls == PF_HIT) out[6] += (double)(t_hi - t_lo) / fmax(1e-30, fabs(t));
        }
        test_leaf(S, node, r, 0.0001, best);
        tbest32 = __double2float_ru(best.t);
      }
      if (sp == 0) break;
      node = stack[--sp];
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// sim_candidates: the whole scheme (closest_candidates + resolve_candidates) against closest_surface on the same
// rays: prim AND t must be identical.  out: 0 rays, 1 mismatches, 2 overflows (re-traced exactly by the kernel),
// 3 sum of candidates resolved, 4 node visits (candidate scheme), 5 node visits (exact scheme), 6 rays with 2
// candidates, 7 prefilter tests.
// ------------------------------------------------------------------------------------------------
extern "C" int sim_candidates(void* p, const QRay* rays, long long n, int K, double* out) {
  (void)K;
  const Emu* e = static_cast<Emu*>(p);
  const DScene& S = e->dev;
  for (int k = 0; k < 8; k++) out[k] = 0.;
  for (long long i = 0; i < n; i++) {
    Ray r;
    r.ox = rays[i].ox; r.oy = rays[i].oy; r.oz = rays[i].oz;
    r.dx = rays[i].dx; r.dy = rays[i].dy; r.dz = rays[i].dz; r.time = rays[i].time;
    Hit exact;
    hit_reset(exact);
    DStats st = {}, sc = {};
    closest_surface<true>(S, r, 0.0001, exact, &st);
    Cands C;
    const bool ok = closest_candidates<true>(S, r, S.scene_mag, C, &sc);
    out[0]++;
    out[4] += (double)sc.node_visits;
    out[5] += (double)st.node_visits;
    out[7] += (double)sc.prim_tests;
    if (!ok) { out[2]++; continue; }
    Hit best;
    int r0, r1;
    cands_record(C, r0, r1);
    resolve_candidates<true>(S, r0, r1, r, 0.0001, best);
    out[3] += (C.c0 < 0) + (C.c1 < 0);
    out[6] += (C.c1 < 0);
    if (best.prim != exact.prim || best.t != exact.t) out[1]++;
  }
  return 0;
}
