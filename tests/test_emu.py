"""CPU tests of the DEVICE LOGIC: csrc/rtb_device.cuh + csrc/flatten.cpp compiled for the host
(tests/emu, a development aid -- not a fallback, never shipped) against the oracle.  They catch
kernel-logic bugs in a container without a GPU; the real parity tests are tests/test_gpu_*.py."""
import numpy as np
import pytest

from oracle import orc
from surely_raytracing_b200 import capi
from surely_raytracing_b200.scenes import BuiltScene
from tests import util
from tests.emu.emu_lib import EmuScene


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4", "c5", "box_city"])
def test_first_hit_parity(cfg):
    b = BuiltScene(cfg, width=128, spp=4)
    o = orc.OracleScene(b, use_bvh=False)
    e = EmuScene(b)
    rays = o.camera_rays()
    ho, he = o.trace(rays), e.trace(rays)
    mism, t_rel, dn, duv = util.hit_errors(ho, he)
    assert mism == 0 and t_rel < 1e-9 and dn < 1e-9 and duv < 1e-9, (mism, t_rel, dn, duv)
    hb = e.trace(rays, capi.RTB_TRACE_BRUTE_FORCE)
    assert (hb["prim"] == he["prim"]).all() and np.array_equal(hb["t"], he["t"])
    sec = util.secondary_rays(ho, np.random.default_rng(7), n_max=6000)
    so, se = o.trace(sec), e.trace(sec)
    mism, t_rel, dn, duv = util.hit_errors(so, se)
    assert mism == 0 and t_rel < 1e-7 and dn < 1e-7 and duv < 1e-7, (mism, t_rel, dn, duv)
    sb = e.trace(sec, capi.RTB_TRACE_BRUTE_FORCE)
    assert (sb["prim"] == se["prim"]).all()


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4", "c5", "earth", "box_city"])
def test_candidate_scheme_finds_the_exact_closest_hit(cfg):
    """The wavefront traversal only CLASSIFIES leaf primitives (conservative fp32 test with error bounds:
    certain miss / certain hit within [t_lo, t_hi] / unsure) and the exact f64 reference-order test runs on the
    <= 2 survivors (rtb_device.cuh: prefilter_*, cands_add, resolve_candidates).  Its closest hit must be the
    exact traversal's, prim AND t bit for bit -- on pixel-centre rays with f64 and with fp32-rounded directions
    (the Cornell diagonal pixels run exactly along a quad edge), secondary rays leaving surfaces (t_min, own
    primitive at t ~ 0), rays grazing their own surface, axis-parallel and far-away rays."""
    b = BuiltScene(cfg, width=160 if cfg != "c4" else 200, spp=4)
    e = EmuScene(b)
    rng = np.random.default_rng(3)
    rays = orc.OracleScene(b).camera_rays()
    he = e.trace(rays)
    r32 = rays.copy()
    r32["direction"] = r32["direction"].astype(np.float32)
    sec = util.secondary_rays(he, rng, n_max=20000)
    sec["direction"] = sec["direction"].astype(np.float32)
    ok = he["prim"] >= 0
    graze = sec.copy()
    idx = rng.integers(0, int(ok.sum()), len(graze))
    nrm = he["normal"][ok][idx]
    graze["origin"] = he["p"][ok][idx]
    tang = np.cross(nrm, rng.normal(size=(len(graze), 3)))
    tang /= np.linalg.norm(tang, axis=1, keepdims=True) + 1e-300
    graze["direction"] = (tang + nrm * rng.choice([1e-3, 1e-5, 1e-7, 0.0, -1e-6], size=(len(graze), 1))).astype(np.float32)
    wild = np.zeros(40000, dtype=capi.RAY_DTYPE)
    M = 1.2 * max(np.abs(rays["origin"]).max(), 600.0)
    wild["origin"] = rng.uniform(-M, M, (len(wild), 3))
    d = rng.normal(size=(len(wild), 3))
    d[:4000, rng.integers(0, 3)] = 0.0
    d[4000:8000, :2] = 0.0
    wild["direction"] = d.astype(np.float32)
    wild["origin"][8000:10000] *= 50.0
    wild["time"] = rng.uniform(0, 1, len(wild))
    wild["t_min"] = 1e-4
    n_two = 0
    for name, rr in (("f64 pixel centres", rays), ("fp32 pixel centres", r32), ("secondary", sec), ("grazing", graze), ("wild", wild)):
        rr = rr.copy()
        rr["time"] = rr["time"].astype(np.float32)           # the queues carry fp32 time
        exact = e.trace(rr)
        cand, st = e.trace_candidates(rr)
        same_t = (exact["t"] == cand["t"]) | (np.isinf(exact["t"]) & np.isinf(cand["t"]))
        assert (exact["prim"] == cand["prim"]).all() and same_t.all(), (name, int((exact["prim"] != cand["prim"]).sum()))
        assert np.array_equal(exact["p"], cand["p"]) and np.array_equal(exact["normal"], cand["normal"])
        if name in ("f64 pixel centres", "secondary"):
            # the exact re-trace stays the rare path (box_city has THREE coplanar faces at y = 0 -- tower bottoms, the slab's top,
            # the glass box's bottom: rays heading down through them overflow the two slots by construction)
            assert st["overflows"] < (0.05 if cfg == "box_city" else 0.005) * len(rr), (name, st)
            assert st["resolved"] < 1.05 * len(rr)                    # <= ~1 exact test per ray instead of ~1.4-2
        n_two += st["two_candidates"]
    assert n_two > 0                                                  # the two-slot path is exercised


@pytest.mark.parametrize("cfg,variant", util.CONFIG_VARIANTS + [("furnace", 0)])
def test_keyed_samples_match_path_by_path(cfg, variant):
    """Same Philox slots and sampling maps in oracle (f64) and device code (fp32 shading): the
    per-pixel sums agree except for the rare path that crosses a discontinuity."""
    b = BuiltScene(cfg, width=48, spp=9, variant=variant)
    o, e = orc.OracleScene(b), EmuScene(b)
    so, _ = o.render(sampler=orc.SAMPLER_KEYED)
    se, st = e.render()
    rel = np.abs(so - se).max(axis=2) / (np.abs(so).max(axis=2) + 1e-3)
    assert (rel > 2e-3).mean() < 0.01, (rel > 2e-3).mean()
    assert abs(so.mean() - se.mean()) < 2e-3 * so.mean()
    assert st["nonfinite_samples"] == 0


def test_medium_intervals_and_textures_and_light_pdf(monkeypatch):
    b = BuiltScene("c4", width=64, spp=4, variant=1)
    o, e = orc.OracleScene(b), EmuScene(b)
    rays = o.camera_rays()
    for m in range(2):
        a0, a1 = o.medium_interval(m, rays)
        b0, b1 = e.medium_interval(m, rays)
        assert np.array_equal(np.isnan(a0), np.isnan(b0))
        ok = ~np.isnan(a0)
        assert np.allclose(a0[ok], b0[ok], rtol=1e-10) and np.allclose(a1[ok], b1[ok], rtol=1e-10)
    # box-bounded media (c3): one scan of the six quads serves both boundary probes (rtb_device.cuh, medium_interval);
    # camera rays, rays started inside / near the boxes in random directions, and rays aimed at box edges and
    # corners (two or three faces hit within 1e-4: the second probe must then run as written)
    b3 = BuiltScene("c3", width=64, spp=4)
    o3, e3 = orc.OracleScene(b3), EmuScene(b3)
    cam = o3.camera_rays()
    rng3 = np.random.default_rng(21)
    inside = cam[:4000].copy()
    inside["origin"] = rng3.uniform([100., -20., 40.], [460., 350., 480.], (4000, 3))
    inside["direction"] = rng3.normal(size=(4000, 3))
    th = np.radians(15.0)   # box 1 of cornell_smoke: 165 x 330 x 165, rotate_y(15 deg), translate (265, 0, 295)
    corners = np.array([[x * np.cos(th) + z * np.sin(th) + 265., y, -x * np.sin(th) + z * np.cos(th) + 295.]
                        for x in (0., 165.) for y in (0., 330.) for z in (0., 165.)])
    edge_pts = np.concatenate([corners, 0.5 * (corners[:, None] + corners[None, :]).reshape(-1, 3)])
    aimed = cam[:len(edge_pts)].copy()
    aimed["direction"] = edge_pts - aimed["origin"]
    for rays3 in (cam, inside):
        for m in range(2):
            a0, a1 = o3.medium_interval(m, rays3)
            b0, b1 = e3.medium_interval(m, rays3)
            assert np.array_equal(np.isnan(a0), np.isnan(b0))
            ok = ~np.isnan(a0)
            assert np.allclose(a0[ok], b0[ok], rtol=1e-10, atol=1e-12) and np.allclose(a1[ok], b1[ok], rtol=1e-10, atol=1e-12)
    # the single scan against the two probes as written (RTB_FLAG_NO_BOX_SCAN): bit for bit, on
    # every ray set -- exact edges included, where the oracle itself may differ in the last bit because the
    # reference tests the box in its own rotated frame and the device in world space (instances)
    e3g = EmuScene(BuiltScene("c3", width=64, spp=4, flags=capi.RTB_FLAG_NO_BOX_SCAN))
    from tests.emu.emu_lib import load as _load
    assert _load().emu_medium_flags(e3._h, 0) & 0x400 and not (_load().emu_medium_flags(e3g._h, 0) & 0x600)
    for rays3 in (cam, inside, aimed):
        for m in range(2):
            b0, b1 = e3.medium_interval(m, rays3)
            g0, g1 = e3g.medium_interval(m, rays3)
            assert np.array_equal(b0, g0, equal_nan=True) and np.array_equal(b1, g1, equal_nan=True)
    assert (~np.isnan(e3.medium_interval(0, aimed)[0])).sum() >= 10
    assert (~np.isnan(o3.medium_interval(0, inside)[0])).sum() > 500
    rng = np.random.default_rng(3)
    n_tex = b.desc.contents.n_textures
    uvp = np.hstack([rng.uniform(-0.2, 1.2, (4000, 2)), rng.uniform(-300, 600, (4000, 3))])
    for t in range(n_tex):
        co, ce = o.eval_texture(t, uvp), e.eval_texture(t, uvp)
        assert np.abs(co - ce).max() < 2e-4, (t, np.abs(co - ce).max())
    od = np.hstack([rng.uniform(100, 450, (4000, 3)), rng.normal(size=(4000, 3))])
    od[:, 4] = np.abs(od[:, 4])
    po, pe = o.eval_light_pdf(od), e.eval_light_pdf(od)
    assert np.allclose(po, pe, rtol=1e-9, atol=0)


@pytest.mark.parametrize("name", ["scene_three_spheres", "two_spheres", "earth", "two_perlin_spheres", "quads", "simple_light"])
def test_remaining_main_rs_scenes(name):
    """The other scene functions of reference src/main.rs: hollow glass via a negative
    radius, checker / image / Perlin textures, emissive sphere + quad."""
    b = BuiltScene(name, width=64, spp=9, variant=1 if name == "simple_light" else 0)
    o, e = orc.OracleScene(b, use_bvh=False), EmuScene(b)
    rays = o.camera_rays()
    mism, t_rel, dn, duv = util.hit_errors(o.trace(rays), e.trace(rays))
    assert mism == 0 and t_rel < 1e-9 and dn < 1e-9 and duv < 1e-9, (mism, t_rel, dn, duv)
    so, _ = o.render(sampler=orc.SAMPLER_KEYED)
    se, st = e.render()
    rel = np.abs(so - se).max(axis=2) / (np.abs(so).max(axis=2) + 1e-3)
    assert (rel > 2e-3).mean() < 0.02 and abs(so.mean() - se.mean()) < 3e-3 * so.mean(), ((rel > 2e-3).mean(), so.mean(), se.mean())


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4", "c5"])
def test_alternative_trees_find_the_same_closest_hits(cfg):
    """The opt-in forms of the tree the wavefront extend can traverse -- 16-bit quantised 32-byte nodes and the
    collapsed BVH4 (csrc/flatten.cpp) -- are conservative: same closest hit (prim AND t, bit for bit) as the
    fp32 BVH2 and as a brute-force scan, on camera rays, surface-leaving rays, axis-parallel directions
    and origins far outside the scene (the quantised form falls back to an unculled walk there)."""
    b = BuiltScene(cfg, width=96, spp=4)
    e = EmuScene(b)
    rays = e_rays = orc.OracleScene(b, use_bvh=False).camera_rays()
    sec = util.secondary_rays(e.trace(rays), np.random.default_rng(11), n_max=8000)
    sec["direction"] = sec["direction"].astype(np.float32)          # queue directions are fp32-valued
    rng = np.random.default_rng(5)
    adv = sec[rng.integers(0, len(sec), 4000)].copy()
    adv["direction"][:1000, 0] = 0.0
    adv["direction"][1000:2000, 1] = 0.0
    adv["origin"][2000:3000] *= 40.0
    adv["origin"][3000:] += rng.normal(0, 3000, (1000, 3))
    cam = e_rays.copy()
    cam["direction"] = cam["direction"].astype(np.float32)
    allr = np.concatenate([cam, sec, adv])
    res = e.check_trees(allr, brute_every=5)
    for name in ("qnodes", "nodes4"):
        assert res[name]["mismatch_bvh2"] == 0 and res[name]["mismatch_brute"] == 0, (name, res[name])
    assert res["nodes4"]["visits"] < 0.75 * res["nodes4"]["visits_bvh2"]   # the collapse does halve the node steps


SPEC = dict(MEDIA=1, BOXSCAN=2, LIGHTS=4, GENERIC_MEDIA=8, QUAD_UV=16, SPHERE_UV=32, TEXTURES=64)


@pytest.mark.parametrize("cfg,variant,expected", [
    ("c1", 0, SPEC["TEXTURES"]),          # checker ground
    ("c2", 0, SPEC["LIGHTS"]),
    ("c3", 0, SPEC["MEDIA"] | SPEC["BOXSCAN"] | SPEC["GENERIC_MEDIA"]),
    ("c3", 1, SPEC["MEDIA"] | SPEC["BOXSCAN"] | SPEC["GENERIC_MEDIA"] | SPEC["LIGHTS"]),
    ("c4", 0, SPEC["MEDIA"] | SPEC["SPHERE_UV"] | SPEC["TEXTURES"]),
    ("c4", 1, SPEC["MEDIA"] | SPEC["SPHERE_UV"] | SPEC["TEXTURES"] | SPEC["LIGHTS"]),
    ("c5", 0, SPEC["LIGHTS"]),
    ("earth", 0, SPEC["SPHERE_UV"] | SPEC["TEXTURES"]),
])
def test_scene_feature_bits_and_leaf_references(cfg, variant, expected):
    """The flattener's feature bits pick the shade kernel instantiation (a missing bit would compile code the scene
    needs OUT, so they are pinned per config), and every BVH leaf reference carries its primitive's kind."""
    e = EmuScene(BuiltScene(cfg, width=32, spp=4, variant=variant))
    assert e.spec_bits() == expected, (cfg, variant, e.spec_bits())
    assert e.leaf_ref_violations() == 0
    # every non-solid texture of these scenes hangs off a Lambertian surface: the textured classes may be deferred
    assert e.defer_ok() == 1
    from tests.emu.emu_lib import load
    flags = [load().emu_medium_flags(e._h, mi) for mi in range(e.info.n_media)]
    if cfg == "c3":   # Translate(RotateY(make_box)) boundaries are recognised as oriented boxes (tight fp32 cull, inside shortcut)
        assert all(f & 0x200 and f & 0x400 for f in flags), flags
    if cfg == "c4":   # sphere boundaries
        assert all(f & 0x100 and not f & 0x400 for f in flags), flags


def test_leaf_reference_codec():
    from tests.emu.emu_lib import load
    lib = load()
    for first in (0, 1, 3406, (1 << 26) - 1):
        for count in (1, 2, 4, 8):
            for kind in (0, 1, 2, 3):
                assert lib.emu_leaf_roundtrip(first, count, kind) == 1


def test_multi_primitive_leaves_on_the_host_build():
    """RTB_FLAG_BVH_LEAF4: leaves of several primitives go through the generic leaf loop (prim_info read per test),
    and the candidate scheme then carries whole leaves as candidates."""
    b = BuiltScene("c4", width=96, spp=4, flags=capi.RTB_FLAG_BVH_LEAF4)
    e = EmuScene(b)
    assert e.leaf_ref_violations() == 0
    assert e.spec_bits() & 128                                             # SPEC_MULTI_LEAF
    rays = orc.OracleScene(b, use_bvh=False).camera_rays()
    he, hb = e.trace(rays), e.trace(rays, capi.RTB_TRACE_BRUTE_FORCE)
    assert (hb["prim"] == he["prim"]).all() and np.array_equal(hb["t"], he["t"])
    hc, st = e.trace_candidates(rays)
    assert (hc["prim"] == he["prim"]).all() and np.array_equal(hc["t"], he["t"]) and st["overflows"] < 0.01 * len(rays)
    b1 = BuiltScene("c4", width=96, spp=4, flags=capi.RTB_FLAG_NO_BOX_LEAVES)
    e1 = EmuScene(b1)
    assert e.info.n_bvh_nodes < 0.7 * e1.info.n_bvh_nodes
    h1 = e1.trace(rays)
    assert (h1["prim"] == he["prim"]).all() and np.array_equal(h1["t"], he["t"])


def test_axis_aligned_boxes_become_one_leaf():
    """Default build: each of the 400 make_box ground boxes of c4 (src/main.rs:439-463) is ONE leaf whose slab test
    names the face; the tree shrinks by the 5 x 400 inner nodes and every hit -- exact arm, candidate arm, brute force,
    the six-leaf build -- stays the same record."""
    b = BuiltScene("c4", width=96, spp=4)
    e = EmuScene(b)
    e1 = EmuScene(BuiltScene("c4", width=96, spp=4, flags=capi.RTB_FLAG_NO_BOX_LEAVES))
    assert e.leaf_ref_violations() == 0 and e1.leaf_ref_violations() == 0
    assert e1.info.n_bvh_nodes - e.info.n_bvh_nodes == 5 * 400
    rays = orc.OracleScene(b, use_bvh=False).camera_rays()
    he, h1, hb = e.trace(rays), e1.trace(rays), e.trace(rays, capi.RTB_TRACE_BRUTE_FORCE)
    hc, st = e.trace_candidates(rays)
    for h in (h1, hb, hc):
        assert (h["prim"] == he["prim"]).all() and np.array_equal(h["t"], he["t"]) and np.array_equal(h["normal"], he["normal"])
    assert st["overflows"] < 0.005 * len(rays)
    # the rotated boxes of c3 (RotateY) are not axis-aligned in world space: they stay six leaves
    c3 = EmuScene(BuiltScene("c3", width=32, spp=4)); c3n = EmuScene(BuiltScene("c3", width=32, spp=4, flags=capi.RTB_FLAG_NO_BOX_LEAVES))
    assert c3.info.n_bvh_nodes == c3n.info.n_bvh_nodes


def test_top_levels_of_the_tree_come_first_in_memory():
    """The flattener numbers the top 7 levels of the BVH breadth-first (what the shared-memory arm of the extend
    kernel stages) and the subtrees below them depth-first: children of the first nodes are themselves early."""
    import ctypes as C
    from tests.emu.emu_lib import load
    lib = load()
    e = EmuScene(BuiltScene("c4", width=32, spp=4))
    lib.emu_node_children.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    kids = np.zeros(2, dtype=np.int32)
    top, frontier = [], [(0, 0)]
    while frontier:
        node, d = frontier.pop(0)
        top.append(node)
        lib.emu_node_children(e._h, node, kids.ctypes.data_as(C.c_void_p))
        frontier += [(int(k), d + 1) for k in kids if k >= 0 and d + 1 < 7]
    assert 32 < len(top) <= 127 and top == list(range(len(top)))          # breadth-first, a prefix of the array
