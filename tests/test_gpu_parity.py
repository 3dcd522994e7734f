"""GPU parity tests proper: librtb200.so (through the C ABI) against the oracle on the same inputs.
Bars (BASELINE.json north_star): first-hit ids bit-exact; t / normal / uv within 1e-5 relative (we hold
1e-9: primitive decisions are f64 on the device); images inside a variance-aware RMSE + bias bound."""
import numpy as np
import pytest

from oracle import orc
from surely_raytracing_b200 import Scene, capi
from surely_raytracing_b200.scenes import BuiltScene
from tests import util

pytestmark = pytest.mark.gpu

PIPELINES = [capi.PIPELINE_MEGAKERNEL, capi.PIPELINE_WAVEFRONT]


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4", "c5"])
def test_first_hit_parity_full_resolution(cfg):
    """pixel-centre primary rays of the config at its BASELINE resolution + seeded secondary rays."""
    b = BuiltScene(cfg, spp=4)
    o = orc.OracleScene(b, use_bvh=False if cfg != "c4" else True)
    g = Scene(b)
    rays = g.camera_rays()
    assert np.array_equal(rays["direction"], o.camera_rays()["direction"])
    if cfg == "c4":
        rays = rays[:: 4]  # 160k rays keep the linear-list-free oracle inside a few seconds
    ho, hg = o.trace(rays), g.trace(rays)
    mism, t_rel, dn, duv = util.hit_errors(ho, hg)
    assert mism == 0, f"{mism} first-hit id mismatches"
    assert t_rel < 1e-9 and dn < 1e-9 and duv < 1e-9, (t_rel, dn, duv)
    assert (ho["front_face"] == hg["front_face"]).all() and (ho["material"] == hg["material"]).all()
    hb = g.trace(rays, capi.RTB_TRACE_BRUTE_FORCE)
    assert (hb["prim"] == hg["prim"]).all() and np.array_equal(hb["t"], hg["t"]), "BVH cull is not conservative"
    o.set_use_bvh(False)
    sec = util.secondary_rays(ho, np.random.default_rng(11), n_max=20000 if cfg != "c4" else 4000)
    so, sg = o.trace(sec), g.trace(sec)
    mism, t_rel, dn, duv = util.hit_errors(so, sg)
    assert mism == 0 and t_rel < 1e-7 and dn < 1e-7 and duv < 1e-7, (mism, t_rel, dn, duv)
    sb = g.trace(sec, capi.RTB_TRACE_BRUTE_FORCE)
    assert (sb["prim"] == sg["prim"]).all()


def _same_hits(a, b):
    same_t = (a["t"] == b["t"]) | (np.isinf(a["t"]) & np.isinf(b["t"]))
    return bool((a["prim"] == b["prim"]).all() and same_t.all() and np.array_equal(a["p"], b["p"]) and
                np.array_equal(a["normal"], b["normal"]) and (a["front_face"] == b["front_face"]).all())


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4", "c5", "box_city"])
def test_first_hit_parity_through_the_kernels_rtb_render_runs(cfg):
    """RTB_TRACE_WAVEFRONT: the rays travel as queue records through k_wf_extend (conservative fp32 classification,
    <= 2 candidates), k_wf_extend_exact (overflows) and the exact resolution of the shade stage -- the SHIPPED traversal,
    not the harness kernel.  Pixel-centre rays at the config's full resolution travel as PRIMARY records (f64
    directions, what get_ray produces): ids bit-exact against the oracle, t / normal / uv to 1e-9.  Seeded secondary rays
    travel as SECONDARY records (directions rounded to fp32, as every scattered ray of the pipeline is) and are held to
    the oracle on the same rounded directions.  The round-1 arm (f64 tests inside the traversal, RTB_OPT_EXACT_LEAVES)
    must give the very same records, and so must the harness kernel."""
    b = BuiltScene(cfg, spp=4)
    o = orc.OracleScene(b, use_bvh=False if cfg != "c4" else True)
    g = Scene(b)
    rays = g.camera_rays()
    full = rays
    if cfg == "c4":
        rays = rays[:: 4]
    ho = o.trace(rays)
    hw = g.trace(rays, capi.RTB_TRACE_WAVEFRONT)
    st = g.render_stats()
    mism, t_rel, dn, duv = util.hit_errors(ho, hw)
    assert mism == 0, f"{mism} first-hit id mismatches through k_wf_extend"
    assert t_rel < 1e-9 and dn < 1e-9 and duv < 1e-9, (t_rel, dn, duv)
    assert (ho["front_face"] == hw["front_face"]).all() and (ho["material"] == hw["material"]).all()
    assert st["overflow_rays"] < 0.005 * len(rays) and st["exact_tests"] < 1.05 * len(rays), st
    # all pixel centres: pipeline kernels == harness kernel == exact-leaves arm, record for record
    hh = g.trace(full)
    hw_full = g.trace(full, capi.RTB_TRACE_WAVEFRONT)
    assert _same_hits(hh, hw_full)
    g.set_option(capi.OPT_EXACT_LEAVES, 1)
    assert _same_hits(hh, g.trace(full, capi.RTB_TRACE_WAVEFRONT))
    g.set_option(capi.OPT_EXACT_LEAVES, 0)
    g.set_option(capi.OPT_SMEM_TOP, 1)
    assert _same_hits(hh, g.trace(full, capi.RTB_TRACE_WAVEFRONT))
    g.set_option(capi.OPT_SMEM_TOP, 0)
    # what carrying bounce-0 directions in fp32 (round 1) would have cost: ids that differ from the f64 reference rays
    r32 = full.copy()
    r32["direction"] = r32["direction"].astype(np.float32)
    h32 = g.trace(r32, capi.RTB_TRACE_WAVEFRONT | capi.RTB_TRACE_SECONDARY)
    print(f"{cfg}: pixel-centre ids that change when the direction is rounded to fp32: {int((h32['prim'] != hh['prim']).sum())} of {len(full)}")
    # secondary records: fp32-rounded directions and time, against the oracle on the same rays
    o.set_use_bvh(False)
    sec = util.secondary_rays(ho, np.random.default_rng(11), n_max=20000 if cfg != "c4" else 4000)
    sec["direction"] = sec["direction"].astype(np.float32)
    sec["time"] = sec["time"].astype(np.float32)
    so, sw = o.trace(sec), g.trace(sec, capi.RTB_TRACE_WAVEFRONT | capi.RTB_TRACE_SECONDARY)
    mism, t_rel, dn, duv = util.hit_errors(so, sw)
    assert mism == 0 and t_rel < 1e-7 and dn < 1e-7 and duv < 1e-7, (mism, t_rel, dn, duv)
    assert _same_hits(g.trace(sec), sw)
    with pytest.raises(capi.RtbError):
        bad = sec[:4].copy()
        bad["t_min"] = 1e-3
        g.trace(bad, capi.RTB_TRACE_WAVEFRONT)


def test_candidate_traversal_equals_brute_force_on_random_rays():
    """2M random rays through c4 (axis-parallel ones included) as queue records: the candidate traversal finds the
    brute-force closest hit, prim and t bit for bit."""
    b = BuiltScene("c4", width=64, spp=4)
    g = Scene(b)
    rng = np.random.default_rng(5)
    n = 2_000_000
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["origin"] = rng.uniform([-1200, -50, -1200], [1200, 700, 1200], (n, 3))
    d = rng.normal(size=(n, 3))
    d[: n // 10, rng.integers(0, 3)] = 0.0
    rays["direction"] = d.astype(np.float32)
    rays["time"] = rng.uniform(0, 1, n).astype(np.float32)
    rays["t_min"] = 1e-4
    hb = g.trace(rays, capi.RTB_TRACE_BRUTE_FORCE)
    for flags in (capi.RTB_TRACE_WAVEFRONT, capi.RTB_TRACE_WAVEFRONT | capi.RTB_TRACE_SECONDARY):
        hw = g.trace(rays, flags)
        assert (hb["prim"] == hw["prim"]).all() and np.array_equal(hb["t"], hw["t"])
    assert g.render_stats()["overflow_rays"] < 0.002 * n


def test_device_philox_known_answers():
    """Random123's Philox4x32-10 known-answer vectors on the DEVICE copy of the generator (rtb_device.cuh), and the
    device against the oracle's copy on random counters."""
    g = Scene(BuiltScene("c2", width=16, spp=4))
    kat = np.array([[0, 0, 0, 0, 0, 0], [0xFFFFFFFF] * 6,
                    [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0]], dtype=np.uint32)
    want = np.array([[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8], [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD],
                     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]], dtype=np.uint32)
    assert np.array_equal(g.philox(kat), want)
    rng = np.random.default_rng(1)
    ck = rng.integers(0, 2 ** 32, (5000, 6), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(g.philox(ck), np.array([orc.philox(c[:4], c[4:]) for c in ck], dtype=np.uint32))


def test_bvh_cull_is_conservative_on_random_rays():
    """2M random rays through c4: BVH traversal == brute force (ids and t bit-identical)."""
    b = BuiltScene("c4", width=64, spp=4)
    g = Scene(b)
    rng = np.random.default_rng(5)
    n = 2_000_000
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["origin"] = rng.uniform([-1200, -50, -1200], [1200, 700, 1200], (n, 3))
    d = rng.normal(size=(n, 3))
    d[: n // 10, rng.integers(0, 3)] = 0.0          # axis-parallel rays: 1/0 slabs
    rays["direction"] = d
    rays["time"] = rng.uniform(0, 1, n)
    rays["t_min"] = 1e-4
    hb, hg = g.trace(rays, capi.RTB_TRACE_BRUTE_FORCE), g.trace(rays)
    assert (hb["prim"] == hg["prim"]).all() and np.array_equal(hb["t"], hg["t"])


def test_medium_intervals_textures_light_pdf_write_color():
    b = BuiltScene("c4", width=64, spp=4, variant=1)
    o, g = orc.OracleScene(b), Scene(b)
    rays = o.camera_rays()
    for m in range(2):
        a0, a1 = o.medium_interval(m, rays)
        b0, b1 = g.medium_interval(m, rays)
        assert np.array_equal(np.isnan(a0), np.isnan(b0))
        ok = ~np.isnan(a0)
        assert np.allclose(a0[ok], b0[ok], rtol=1e-10) and np.allclose(a1[ok], b1[ok], rtol=1e-10)
    b3 = BuiltScene("c3", width=64, spp=4)
    o3, g3 = orc.OracleScene(b3), Scene(b3)
    r3 = o3.camera_rays()
    for m in range(2):  # boundaries that are rotated+translated boxes of six quads
        a0, a1 = o3.medium_interval(m, r3)
        b0, b1 = g3.medium_interval(m, r3)
        assert np.array_equal(np.isnan(a0), np.isnan(b0)) and (~np.isnan(a0)).sum() > 100
        ok = ~np.isnan(a0)
        assert np.allclose(a0[ok], b0[ok], rtol=1e-10) and np.allclose(a1[ok], b1[ok], rtol=1e-10)
    rng = np.random.default_rng(3)
    uvp = np.hstack([rng.uniform(-0.2, 1.2, (20000, 2)), rng.uniform(-300, 600, (20000, 3))])
    for t in range(b.desc.contents.n_textures):
        co, cg = o.eval_texture(t, uvp), g.eval_texture(t, uvp)
        assert np.abs(co - cg).max() < 2e-4, (t, np.abs(co - cg).max())
    b1_ = BuiltScene("c1", width=64, spp=4)  # checker texture
    o1, g1 = orc.OracleScene(b1_), Scene(b1_)
    uvp1 = np.hstack([rng.uniform(0, 1, (20000, 2)), rng.uniform(-12, 12, (20000, 3))])
    assert np.array_equal(o1.eval_texture(0, uvp1), g1.eval_texture(0, uvp1).astype(np.float32).astype(np.float64)) or \
        np.abs(o1.eval_texture(0, uvp1) - g1.eval_texture(0, uvp1)).max() < 1e-6
    b5 = BuiltScene("c5", width=64, spp=4)
    o5, g5 = orc.OracleScene(b5), Scene(b5)
    od = np.hstack([rng.uniform(50, 500, (50000, 3)), rng.normal(size=(50000, 3))])
    po, pg = o5.eval_light_pdf(od), g5.eval_light_pdf(od)
    finite = np.isfinite(po)
    assert np.array_equal(np.isfinite(pg), finite) and (po[finite] > 0).sum() > 1000
    assert np.allclose(po[finite], pg[finite], rtol=1e-9, atol=0)
    # output stage: byte-identical to write_color (src/color.rs:8-33)
    px = np.abs(rng.normal(size=(100000, 3))) * rng.choice([0.001, 0.1, 1.0, 30.0], size=(100000, 1)) * 7
    px[5] = np.nan
    for exposure in (0.0, 1.3):
        assert np.array_equal(orc.write_color(px, 7.0, exposure), g.write_color(px, 7.0, exposure))


@pytest.mark.parametrize("pipeline", PIPELINES)
@pytest.mark.parametrize("cfg,variant", util.CONFIG_VARIANTS + [("furnace", 0)])
def test_keyed_samples_match_path_by_path(cfg, variant, pipeline):
    """Oracle in KEYED mode draws the same Philox slots through the same sampling maps (in f64): per-pixel
    sums of a few strata agree except for the rare path that crosses a discontinuity."""
    b = BuiltScene(cfg, width=96, spp=16, variant=variant)
    o, g = orc.OracleScene(b), Scene(b)
    so, _ = o.render(sampler=orc.SAMPLER_KEYED)
    sg, st = g.render(pipeline=pipeline, collect_stats=True)
    rel = np.abs(so - sg).max(axis=2) / (np.abs(so).max(axis=2) + 1e-3)
    assert (rel > 2e-3).mean() < 0.01, (rel > 2e-3).mean()
    assert abs(so.mean() - sg.mean()) < 2e-3 * so.mean()
    assert st["nonfinite_samples"] == 0 and st["paths"] == so.shape[0] * so.shape[1] * 16


@pytest.mark.parametrize("pipeline", PIPELINES)
@pytest.mark.parametrize("name", ["oracle_c1", "oracle_c2", "oracle_c3", "oracle_c3_lights", "oracle_c4",
                                   "oracle_c4_lights", "oracle_c5"])
def test_image_statistics_against_the_committed_oracle_render(name, pipeline):
    """Variance-aware acceptance of the GPU image against the oracle's reference-sampler
    (rejection loops, sequential f64 stream) render committed under tests/golden/: independent RNG,
    independent code.  RMSE <= 1.15 sigma, |bias| <= 4 sigma/sqrt(n_px), per channel, radiance in [0,10]."""
    gold = np.load(util.GOLDEN / f"{name}.npz")
    cfg = name.split("_")[1]
    b = BuiltScene(cfg, width=int(gold["width"]), spp=int(gold["spp"]), variant=int(gold["variant"]))
    g = Scene(b)
    n = g.info.spp_used
    assert n == int(gold["spp"]) and (g.info.image_height, g.info.image_width) == gold["mean"].shape[:2]
    sg, st = g.render(pipeline=pipeline)
    ok, rep = util.image_acceptance(sg / n, n, gold["mean"].astype(np.float64), int(gold["spp"]), gold["var"].astype(np.float64))
    img_g = orc.write_color(sg, n)
    img_o = orc.write_color(gold["mean"].astype(np.float64), 1.0)
    print(name, rep, "PSNR(8-bit sRGB) %.2f dB" % util.psnr8(img_g, img_o))
    assert ok, rep


@pytest.mark.parametrize("pipeline", PIPELINES)
def test_bit_reproducible_and_additive_over_sample_ranges(pipeline):
    """"Every render reproducible" (north star): the accumulation buffer holds 64-bit fixed-point sums added with integer
    atomics, so the image does not depend on the order paths finish in -- same seed -> same BITS, run to run, and for
    any split of the stratum range over calls (what ranks do), in both pipelines."""
    b = BuiltScene("c5", width=128, spp=64)
    g = Scene(b)
    a, _ = g.render(pipeline=pipeline)
    a2, _ = Scene(b).render(pipeline=pipeline)
    assert np.array_equal(a, a2)
    parts = np.zeros_like(a)
    acc = np.zeros_like(a)
    for lo, hi in ((0, 10), (10, 33), (33, 64)):
        g.render(lo, hi, pipeline=pipeline, out=parts)   # accumulates INTO `parts` (Q24)
    # (the three host-side f64 additions of `parts` round; the device sums themselves are exact -- checked through
    #  rtb_render_device in test_device_buffers_of_disjoint_ranges_add_up_exactly)
    assert np.allclose(parts, a, rtol=1e-14, atol=0)
    g2 = Scene(BuiltScene("c5", width=128, spp=64, seed=99))
    c, _ = g2.render(pipeline=pipeline)
    assert not np.array_equal(a, c) and abs(a.mean() - c.mean()) < 0.05 * a.mean()


def test_device_buffers_of_disjoint_ranges_add_up_exactly():
    """rtb_render_device accumulates INTO a caller-owned u64 x 4 buffer: one call over [0, 64) and three calls over a
    split of it leave the same bits (what makes the multi-GPU int64 reduce split-invariant), whichever pipeline
    renders which part; counts = strata per pixel."""
    import torch
    b = BuiltScene("c3", width=96, spp=64, variant=1)
    g = Scene(b)
    h, w = g.info.image_height, g.info.image_width
    one = torch.zeros((h, w, 4), dtype=torch.int64, device="cuda")
    g.render_device(one.data_ptr(), 0, 64, pipeline=capi.PIPELINE_WAVEFRONT)
    torch.cuda.synchronize()
    split = torch.zeros_like(one)
    for (lo, hi), pipeline in (((0, 7), capi.PIPELINE_WAVEFRONT), ((7, 40), capi.PIPELINE_WAVEFRONT), ((40, 64), capi.PIPELINE_WAVEFRONT)):
        g.render_device(split.data_ptr(), lo, hi, pipeline=pipeline)
    torch.cuda.synchronize()
    assert torch.equal(one, split) and int(one[..., 3].min()) == int(one[..., 3].max()) == 64
    px = g.accum_to_pixels(one.data_ptr())
    ref, _ = g.render(pipeline=capi.PIPELINE_WAVEFRONT)
    assert np.array_equal(px, ref)
    assert np.array_equal(px, one[..., :3].cpu().numpy().astype(np.float64) / capi.ACCUM_SCALE)


@pytest.mark.parametrize("cfg,variant", [("c4", 0), ("c4", 1), ("c1", 0), ("c3", 0), ("c5", 0)])
def test_candidate_scheme_renders_the_very_same_image(cfg, variant):
    """The traversal decides nothing that could change a result: with the f64 tests back inside the traversal
    (RTB_OPT_EXACT_LEAVES: the round-1 kernel) or the tree's top staged in shared memory (RTB_OPT_SMEM_TOP) every ray
    finds the same hit, every path is the same path, and the image is the same image BIT FOR BIT."""
    b = BuiltScene(cfg, width=200, spp=16, variant=variant)
    ref, st = Scene(b).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    exact, st_e = Scene(b).set_option(capi.OPT_EXACT_LEAVES, 1).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert np.array_equal(ref, exact)
    assert st["segments"] == st_e["segments"] and st_e["overflow_rays"] == 0
    assert st["exact_tests"] < 0.75 * st_e["prim_tests"]            # fewer f64 tests, and those at dense lanes
    assert st["overflow_rays"] < 0.005 * st["segments"]
    smem, _ = Scene(b).set_option(capi.OPT_SMEM_TOP, 1).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert np.array_equal(ref, smem)
    # ... and in the uncounted (scene-specialised) instantiations that production calls run
    ref2, _ = Scene(b).render(pipeline=capi.PIPELINE_WAVEFRONT)
    exact2, _ = Scene(b).set_option(capi.OPT_EXACT_LEAVES, 1).render(pipeline=capi.PIPELINE_WAVEFRONT)
    assert np.array_equal(ref2, exact2)


@pytest.mark.parametrize("flag", [capi.RTB_FLAG_BVH4, capi.RTB_FLAG_QNODES])
@pytest.mark.parametrize("cfg", ["c4", "c2"])
def test_opt_in_tree_forms_give_the_same_image(cfg, flag):
    """The extend kernel's other node formats (collapsed BVH4, 16-bit quantised nodes; chosen when the scene
    is created) only change the cull: same closest hits, hence the same image bit for bit, and strictly fewer
    node visits for the BVH4."""
    ref, st_ref = Scene(BuiltScene(cfg, width=160, spp=16)).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    alt, st_alt = Scene(BuiltScene(cfg, width=160, spp=16, flags=flag)).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert np.array_equal(alt, ref)
    assert st_alt["segments"] == st_ref["segments"] and st_alt["prim_tests"] <= 1.1 * st_ref["prim_tests"]
    if flag == capi.RTB_FLAG_BVH4:
        assert st_alt["node_visits"] < 0.75 * st_ref["node_visits"]


def test_box_leaves_give_the_same_image():
    """Default build: an axis-aligned make_box is one leaf whose slab test names the candidate face.  Against the build
    with one leaf per quad (RTB_FLAG_NO_BOX_LEAVES): the same image bit for bit, on both extend arms."""
    for cfg in ("c4", "c3", "box_city"):   # box_city (host/scenes.cpp): touching, nested and translated boxes, the camera inside one
        ref, _ = Scene(BuiltScene(cfg, width=160, spp=16, flags=capi.RTB_FLAG_NO_BOX_LEAVES)).render(pipeline=capi.PIPELINE_WAVEFRONT)
        g = Scene(BuiltScene(cfg, width=160, spp=16))
        assert ref.max() > 0 and np.array_equal(g.render(pipeline=capi.PIPELINE_WAVEFRONT)[0], ref), cfg
        g.set_option(capi.OPT_EXACT_LEAVES, 1)
        assert np.array_equal(g.render(pipeline=capi.PIPELINE_WAVEFRONT)[0], ref), cfg


def test_box_media_face_naming_gives_the_same_image():
    """c3's two smoke boxes: the oriented-box slab test names the entry and exit face and two plane distances replace the
    six-quad scan (rtb_device.cuh, medium_obb / medium_interval).  Against RTB_FLAG_NO_BOX_SCAN -- both boundary probes of
    constant_medium.rs:46-55 as written, over all six quads -- the image is the same bit for bit, on both pipelines."""
    for pipeline in (capi.PIPELINE_WAVEFRONT, capi.PIPELINE_MEGAKERNEL):
        ref, st_ref = Scene(BuiltScene("c3", width=200, spp=36, flags=capi.RTB_FLAG_NO_BOX_SCAN)).render(pipeline=pipeline, collect_stats=True)
        img, st = Scene(BuiltScene("c3", width=200, spp=36)).render(pipeline=pipeline, collect_stats=True)
        assert ref.max() > 0 and np.array_equal(img, ref), pipeline
        assert st["segments"] == st_ref["segments"] and st["medium_probes"] == st_ref["medium_probes"]


def test_multi_primitive_leaves_give_the_same_image():
    """RTB_FLAG_BVH_LEAF4 (a tuning arm of the builder) makes leaves of several primitives: the extend kernel then
    runs its generic-leaf instantiation and whole leaves travel as candidates.  Same closest hits, same image; fewer
    nodes, more primitive tests."""
    ref, st_ref = Scene(BuiltScene("c4", width=160, spp=16)).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    g = Scene(BuiltScene("c4", width=160, spp=16, flags=capi.RTB_FLAG_BVH_LEAF4))
    alt, st = g.render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert st["segments"] == st_ref["segments"] and st["prim_tests"] > 1.5 * st_ref["prim_tests"]
    _, st_six = Scene(BuiltScene("c4", width=160, spp=16, flags=capi.RTB_FLAG_NO_BOX_LEAVES)).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert st["node_visits"] < st_six["node_visits"]   # (against one leaf per quad; the default's box leaves visit fewer still)
    assert st_ref["node_visits"] < 0.9 * st_six["node_visits"] and st_ref["segments"] == st_six["segments"]
    # (another shade instantiation -- the generic one -- shades these paths: last-bit FMA differences part a few of them)
    rel = np.abs(alt - ref).max(axis=2) / (np.abs(ref).max(axis=2) + 1e-3)
    assert (rel > 1e-3).mean() < 2e-3 and abs(alt.mean() - ref.mean()) < 1e-5 * ref.mean()
    rays = g.camera_rays()[::7]
    hb, hg, hw = g.trace(rays, capi.RTB_TRACE_BRUTE_FORCE), g.trace(rays), g.trace(rays, capi.RTB_TRACE_WAVEFRONT)
    assert (hb["prim"] == hg["prim"]).all() and np.array_equal(hb["t"], hg["t"])
    assert (hb["prim"] == hw["prim"]).all() and np.array_equal(hb["t"], hw["t"])


def test_deferred_textured_classes_give_the_same_image():
    """By default the image / Perlin Lambertian items are shaded by k_wf_shade_rare from a deferred list and the
    main shade kernel carries no texture code; RTB_OPT_NO_DEFER_RARE shades everything in place.  Same rays, same
    hits -- on the scene with both textured spheres, with and without a light list."""
    for variant in (0, 1):
        b = BuiltScene("c4", width=200, spp=16, variant=variant)
        on, st_on = Scene(b).render(pipeline=capi.PIPELINE_WAVEFRONT)
        off, st_off = Scene(b).set_option(capi.OPT_NO_DEFER_RARE, 1).render(pipeline=capi.PIPELINE_WAVEFRONT)
        assert st_on["kernel_launches"] > st_off["kernel_launches"]      # the extra k_wf_shade_rare per iteration
        # The two arms shade through different instantiations of the same fp32 code (other FMA contraction): a path
        # whose sampled direction differs in the last bit may later cross a discontinuity and part ways.  Measured
        # per cause below: the pixels that differ at all, and those beyond 1e-3 relative.
        rel = np.abs(on - off).max(axis=2) / (np.abs(off).max(axis=2) + 1e-3)
        print(f"variant {variant}: pixels that differ {(on != off).any(axis=2).mean():.4f}, beyond 1e-5 {(rel > 1e-5).mean():.5f}, "
              f"beyond 1e-3 {(rel > 1e-3).mean():.5f}, max {rel.max():.2e}, means {on.mean():.9f} {off.mean():.9f}")
        assert (rel > 1e-3).mean() < 2e-3 and abs(on.mean() - off.mean()) < 1e-5 * off.mean(), ((rel > 1e-3).mean(), on.mean(), off.mean())
    # a scene whose only textured material is NOT a plain Lambertian surface must not defer: simple_light's
    # textures are Perlin spheres + solid lights (defers), two_perlin_spheres too; furnace-like scenes have none
    b = BuiltScene("two_perlin_spheres", width=96, spp=9)
    on, _ = Scene(b).render(pipeline=capi.PIPELINE_WAVEFRONT)
    mega, _ = Scene(b).render(pipeline=capi.PIPELINE_MEGAKERNEL)
    rel = np.abs(on - mega).max(axis=2) / (np.abs(mega).max(axis=2) + 1e-3)
    assert (rel > 2e-3).mean() < 0.01 and abs(on.mean() - mega.mean()) < 2e-3 * mega.mean()


@pytest.mark.parametrize("capacity", [1024, 5000, 65536])
def test_small_queues_refill_and_drain_to_the_same_image(capacity):
    """The wavefront queue is topped up every iteration and its launches shrink with the draining tail; a queue
    far smaller than the job (down to the 1024-slot minimum, and a size that is no multiple of a block)
    exercises every refill / partial-block / tail-sizing path.  Philox keys make the image independent of
    the schedule -- bit for bit, now that the accumulation does not depend on the order of the adds."""
    b = BuiltScene("c3", width=96, spp=36, variant=1)
    g = Scene(b).set_option(capi.OPT_FINISH_BELOW, 0)
    ref, st_ref = g.render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    mega, _ = g.render(pipeline=capi.PIPELINE_MEGAKERNEL)
    small, st = Scene(b).set_option(capi.OPT_WF_CAPACITY, capacity).set_option(capi.OPT_FINISH_BELOW, 0).render(
        pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert st["paths"] == st_ref["paths"] and st["segments"] == st_ref["segments"]
    assert st["kernel_launches"] > st_ref["kernel_launches"]
    assert np.array_equal(small, ref) and np.allclose(small, mega, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("cfg", ["c1", "c3", "c4"])
def test_finishing_kernel_ends_the_tail_with_the_same_paths(cfg):
    """Once every path has started and few are left (RTB_OPT_FINISH_BELOW, default 65536), k_wf_finish runs each of them
    to its end instead of ~40 more iterations of launches: same segments; same image up to the last-bit differences
    between the shade instantiations (a few paths part ways), far fewer launches on a depth-50 scene."""
    b = BuiltScene(cfg, width=240, spp=16)
    loop, st_loop = Scene(b).set_option(capi.OPT_FINISH_BELOW, 0).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    fin, st_fin = Scene(b).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert st_fin["kernel_launches"] <= st_loop["kernel_launches"]
    if cfg == "c1":                                                   # depth 50: dozens of nearly empty iterations saved
        assert st_fin["kernel_launches"] < 0.7 * st_loop["kernel_launches"]
    assert abs(st_fin["segments"] - st_loop["segments"]) <= 1e-4 * st_loop["segments"]
    rel = np.abs(fin - loop).max(axis=2) / (np.abs(loop).max(axis=2) + 1e-3)
    assert (rel > 1e-3).mean() < 2e-3 and abs(fin.mean() - loop.mean()) < 1e-5 * loop.mean(), ((rel > 1e-3).mean(), fin.mean(), loop.mean())
    always, st_a = Scene(b).set_option(capi.OPT_FINISH_BELOW, 1 << 24).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    assert abs(st_a["segments"] - st_loop["segments"]) <= 1e-4 * st_loop["segments"]
    rel = np.abs(always - loop).max(axis=2) / (np.abs(loop).max(axis=2) + 1e-3)
    assert (rel > 1e-3).mean() < 5e-3 and abs(always.mean() - loop.mean()) < 1e-4 * loop.mean()


def test_small_calls_take_the_megakernel_and_depth_zero_traces_nothing():
    """RTB_PIPELINE_DEFAULT: below 2^19 paths per call the wavefront is launch-bound and the megakernel renders the
    call (one launch); max_depth 0 returns black like ray_color's `depth <= 0` (render.rs:260-262), max_depth 1 sees
    emitters and background only -- both as the oracle has them."""
    b = BuiltScene("c5", width=64, spp=16)
    g = Scene(b)
    _, st = g.render()
    assert st["kernel_launches"] == 1
    _, st = g.render(pipeline=capi.PIPELINE_WAVEFRONT)
    assert st["kernel_launches"] > 4
    _, st = Scene(b).set_option(capi.OPT_MEGA_BELOW, 0).render()
    assert st["kernel_launches"] > 4
    for cfg in ("c5", "c1"):
        for depth in (0, 1):
            bd = BuiltScene(cfg, width=64, spp=16, depth=depth) if depth else None
            if depth == 0:
                bd = BuiltScene(cfg, width=64, spp=16)
                bd.desc.contents.camera.max_depth = 0
            so, _ = orc.OracleScene(bd).render(sampler=orc.SAMPLER_KEYED)
            for pipeline in PIPELINES:
                sg, stg = Scene(bd).render(pipeline=pipeline)
                assert stg["paths"] == 64 * sg.shape[0] * 16
                if depth == 0:
                    assert not sg.any() and not so.any()
                else:
                    assert np.allclose(sg, so, rtol=1e-5, atol=1e-6), (cfg, depth, pipeline)


def test_sun_light_flag_and_auto_exposure_against_the_oracle():
    """scene_sun_spheres (reference src/main.rs:32-90, scene -2): `suns` is ignored at HEAD (Q23) -- with and without
    the sun records the image is the same -- and RTB_FLAG_SUN_LIGHT restores the commented-out term `background +
    sun_light` (render.rs:300-308) exactly as the oracle restates it.  auto_expose (render.rs:325-339) of the sums and
    the exposed write_color bytes equal the oracle's."""
    b_off = BuiltScene("scene_sun_spheres", width=160, spp=16)
    b_on = BuiltScene("scene_sun_spheres", width=160, spp=16, flags=capi.RTB_FLAG_SUN_LIGHT)
    assert b_on.desc.contents.n_suns == 1
    off, _ = Scene(b_off).render(pipeline=capi.PIPELINE_WAVEFRONT)
    no_suns = BuiltScene("scene_sun_spheres", width=160, spp=16)
    no_suns.desc.contents.n_suns = 0
    assert np.array_equal(off, Scene(no_suns).render(pipeline=capi.PIPELINE_WAVEFRONT)[0])
    so, _ = orc.OracleScene(b_on).render(sampler=orc.SAMPLER_KEYED)
    for pipeline in PIPELINES:
        on, _ = Scene(b_on).render(pipeline=pipeline)
        assert on.mean() > 1.3 * off.mean()
        rel = np.abs(so - on).max(axis=2) / (np.abs(so).max(axis=2) + 1e-3)
        assert (rel > 2e-3).mean() < 0.02 and abs(so.mean() - on.mean()) < 3e-3 * so.mean(), (pipeline, (rel > 2e-3).mean())
    from surely_raytracing_b200 import auto_expose
    e_g, e_o = auto_expose(on, 16), orc.auto_expose(on, 16)
    assert e_g == e_o and 0.05 < e_g < 50 and e_g != 1.0
    g = Scene(b_on)
    assert np.array_equal(g.write_color(on, 16, e_g), orc.write_color(on, 16, e_o))
    assert auto_expose(np.zeros((8, 8, 3)), 4) == 1.0 == orc.auto_expose(np.zeros((8, 8, 3)), 4)


def test_checkpoint_resume_is_bit_identical_and_russian_roulette_is_unbiased(tmp_path):
    """SURVEY 8f-4 extras.  (1) On-disk checkpoints: render [0, 25), save, load into a fresh scene, render [25, 64) ->
    the very bits of one uninterrupted render (the accumulation buffer holds exact integer sums); a checkpoint of another
    seed or size is refused.  (2) RTB_FLAG_RUSSIAN_ROULETTE (opt-in, not in the reference): fewer segments, and the image
    still passes the variance-aware acceptance against the committed oracle render."""
    import torch
    b = BuiltScene("c2", width=120, spp=64)
    g = Scene(b)
    h, w = g.info.image_height, g.info.image_width
    whole = torch.zeros((h, w, 4), dtype=torch.int64, device="cuda")
    g.render_device(whole.data_ptr(), 0, 64, pipeline=capi.PIPELINE_WAVEFRONT)
    part = torch.zeros_like(whole)
    g.render_device(part.data_ptr(), 0, 25, pipeline=capi.PIPELINE_WAVEFRONT)
    torch.cuda.synchronize()
    ck = tmp_path / "c2.rtbck"
    g.checkpoint_save(part.data_ptr(), ck)
    assert ck.stat().st_size == 64 + h * w * 32
    g2 = Scene(b)
    resumed = torch.full_like(whole, 7)
    g2.checkpoint_load(resumed.data_ptr(), ck)
    g2.render_device(resumed.data_ptr(), 25, 64, pipeline=capi.PIPELINE_WAVEFRONT)
    torch.cuda.synchronize()
    assert torch.equal(whole, resumed)
    with pytest.raises(capi.RtbError):
        Scene(BuiltScene("c2", width=120, spp=64, seed=5)).checkpoint_load(resumed.data_ptr(), ck)
    with pytest.raises(capi.RtbError):
        Scene(BuiltScene("c2", width=96, spp=64)).checkpoint_load(resumed.data_ptr(), ck)
    gold = np.load(util.GOLDEN / "oracle_c3.npz")     # lights = empty: paths run to depth 10 unless they meet the light
    n = int(gold["spp"])
    plain, st_p = Scene(BuiltScene("c3", width=120, spp=n)).render(collect_stats=True)
    rr, st_r = Scene(BuiltScene("c3", width=120, spp=n, flags=capi.RTB_FLAG_RUSSIAN_ROULETTE)).render(collect_stats=True)
    assert st_r["segments"] < 0.85 * st_p["segments"]                 # measured 0.80 on c3 (depth 10, roulette from the 4th bounce)
    ok, rep = util.image_acceptance(rr / n, n, gold["mean"].astype(np.float64), n, 1.6 * gold["var"].astype(np.float64))
    assert ok, rep                                        # (roulette adds variance per path: the bound uses 1.6 sigma^2)
    assert abs(rr.mean() - plain.mean()) < 0.01 * plain.mean()


def test_render_multi_is_the_same_image_on_any_number_of_gpus():
    """rtb_render_multi = the reference seam on one box: contiguous slices of the stratum range on n GPUs (one thread,
    scene copy and stream each), ONE NCCL int64 sum-reduce, one D2H.  The image must not depend on n: bit-identical to
    rtb_render on one GPU (runs with every device count the box has, 1 included)."""
    from surely_raytracing_b200 import render_multi
    lib = capi.load_library()
    n_gpu = lib.rtb_device_count()
    b = BuiltScene("c4", width=200, spp=64, variant=1)
    ref, st_ref = Scene(b).render(pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
    for n in sorted({1, 2, n_gpu} & set(range(1, n_gpu + 1))):
        px, st = render_multi(b, n, pipeline=capi.PIPELINE_WAVEFRONT, collect_stats=True)
        assert np.array_equal(px, ref), n
        assert st["paths"] == st_ref["paths"] and st["segments"] == st_ref["segments"]
    out = np.full_like(ref, 1.5)
    render_multi(b, 1, out=out)
    assert np.array_equal(out, ref + 1.5) or np.allclose(out, ref + 1.5, rtol=1e-15)
    with pytest.raises(capi.RtbError):
        render_multi(b, n_gpu + 1)
    with pytest.raises(capi.RtbError):
        render_multi(b, 1, 0, 10 ** 6)
    assert lib.rtb_trim_cache() > 0


def test_flags_iso_pdf_zero_and_full_size_round_trip_property():
    """F3 flag reaches the device; at BASELINE size (c3 600x600) a size-independent property:
    with black-albedo smoke only (density up) the image can only get darker."""
    a = Scene(BuiltScene("c3", width=200, spp=64))
    z = Scene(BuiltScene("c3", width=200, spp=64, flags=capi.RTB_FLAG_ISO_PDF_ZERO))
    sa, _ = a.render()
    sz, _ = z.render()
    assert sz.mean() < sa.mean() * 0.98 and np.isfinite(sa).all() and np.isfinite(sz).all()
    oz = orc.OracleScene(BuiltScene("c3", width=200, spp=64, flags=capi.RTB_FLAG_ISO_PDF_ZERO))
    so, _ = oz.render(sampler=orc.SAMPLER_KEYED)
    assert abs(so.mean() - sz.mean()) < 3e-3 * so.mean()


def test_edge_inputs_empty_world_single_pixel_single_sample():
    """Edge inputs of the hot path: a world with nothing in it (every path returns the background, render.rs:262-270 on a
    miss), a 1 x 1 image, one sample per pixel, an empty stratum range -- on both pipelines, against the oracle."""
    for pipeline in (capi.PIPELINE_WAVEFRONT, capi.PIPELINE_MEGAKERNEL):
        b = BuiltScene("c1", width=24, spp=4)
        d = b.desc.contents
        d.objects[d.world].count = 0                      # HittableList::new() with nothing added
        g = Scene(b)
        assert g.info.n_surface_prims == 0
        img, st = g.render(pipeline=pipeline, collect_stats=True)
        bg = np.array([d.camera.background[0], d.camera.background[1], d.camera.background[2]])
        assert np.allclose(img, 4 * bg, rtol=0, atol=1e-6) and st["segments"] == st["paths"] == img.shape[0] * img.shape[1] * 4
        oimg, _ = orc.OracleScene(b).render(0, 4)
        assert np.allclose(oimg, img, rtol=0, atol=1e-6)   # (the device holds the background in fp32)
        img0, _ = g.render(2, 2, pipeline=pipeline)       # empty range: nothing added, nothing counted
        assert not img0.any()
        rays = g.camera_rays()
        assert (g.trace(rays)["prim"] == -1).all() and (g.trace(rays, capi.RTB_TRACE_WAVEFRONT)["prim"] == -1).all()
        one = BuiltScene("c5", width=1, spp=1)
        g1 = Scene(one)
        assert (g1.info.image_width, g1.info.image_height, g1.info.spp_used) == (1, 1, 1)
        p1, s1 = g1.render(pipeline=pipeline, collect_stats=True)
        assert p1.shape == (1, 1, 3) and s1["paths"] == 1 and np.isfinite(p1).all()
        k1, _ = orc.OracleScene(one).render(0, 1, sampler=orc.SAMPLER_KEYED)   # the same Philox slots: the same single path
        assert np.allclose(k1, p1, rtol=2e-3, atol=1e-5)


def test_concurrent_host_threads_render_their_own_scenes():
    """Two host threads, each with its own scene on the same GPU (ctypes releases the GIL inside the calls): the
    process-wide buffer cache and the per-thread streams must keep them apart -- both images equal the serial ones bit
    for bit, over several rounds of create / render / destroy."""
    import threading
    want = {cfg: Scene(BuiltScene(cfg, width=120, spp=16)).render()[0] for cfg in ("c2", "c4")}
    got, errs = {}, []

    def work(cfg):
        try:
            for _ in range(4):
                s = Scene(BuiltScene(cfg, width=120, spp=16))
                got[cfg] = s.render()[0]
                s.close()
        except Exception as e:   # noqa: BLE001
            errs.append((cfg, repr(e)))

    ts = [threading.Thread(target=work, args=(cfg,)) for cfg in want]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for cfg in want:
        assert np.array_equal(got[cfg], want[cfg]), cfg


def test_host_buffer_accumulates_into_and_rejects_bad_ranges():
    g = Scene(BuiltScene("c2", width=64, spp=16))
    out = np.full((64, 64, 3), 2.0)
    res, _ = g.render(out=out)
    assert res is out and (out >= 2.0).all() and out.mean() > 2.0
    with pytest.raises(capi.RtbError):
        g.render(0, 17)
    with pytest.raises(capi.RtbError):
        g.render(5, 3)


def test_cpp_drop_in_example_writes_the_same_ppm_as_the_python_path(tmp_path):
    """host/example_cornell.cpp = cornell_box written like reference src/main.rs:417-512 on the C++ API
    mirror, rendered by the drop-in render_par_lights: its P3 output must equal write_color of the
    same scene rendered through the Python binding (same seed -> same Philox streams)."""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "surely_raytracing_b200" / "example_cornell"
    assert exe.exists(), "run __graft_entry__.build()"
    r = subprocess.run([str(exe), "64", "16"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    assert lines[0] == "P3" and lines[1] == "64 64" and lines[2] == "255"
    ppm = np.array([[int(v) for v in l.split()] for l in lines[3:3 + 64 * 64]], dtype=np.uint8).reshape(64, 64, 3)
    g = Scene(BuiltScene("c5", width=64, spp=16))
    s, _ = g.render()
    ours = g.write_color(s, g.info.spp_used)
    assert np.array_equal(ppm, ours)                         # same seed, same pipeline choice, order-independent sums


@pytest.mark.parametrize("name", ["scene_three_spheres", "two_spheres", "earth", "two_perlin_spheres", "quads", "simple_light"])
def test_remaining_main_rs_scenes(name):
    """The other scene functions of reference src/main.rs through the C ABI."""
    b = BuiltScene(name, width=160, spp=16, variant=1 if name == "simple_light" else 0)
    o, g = orc.OracleScene(b, use_bvh=False), Scene(b)
    rays = g.camera_rays()
    mism, t_rel, dn, duv = util.hit_errors(o.trace(rays), g.trace(rays))
    assert mism == 0 and t_rel < 1e-9 and dn < 1e-9 and duv < 1e-9, (mism, t_rel, dn, duv)
    so, _ = o.render(sampler=orc.SAMPLER_KEYED)
    for pipeline in PIPELINES:
        sg, st = g.render(pipeline=pipeline)
        rel = np.abs(so - sg).max(axis=2) / (np.abs(so).max(axis=2) + 1e-3)
        assert (rel > 2e-3).mean() < 0.02 and abs(so.mean() - sg.mean()) < 3e-3 * so.mean(), (pipeline, (rel > 2e-3).mean())


@pytest.mark.parametrize("name,cfg,variant", [("oracle_c1", "c1", 0), ("oracle_c2", "c2", 0), ("oracle_c3", "c3", 0),
                                               ("oracle_c4", "c4", 0), ("oracle_c5", "c5", 0)])
def test_full_size_configs_against_the_oracle_statistics(name, cfg, variant):
    """BASELINE.json's configs at their FULL resolution and FULL spp on the GPU (c4: 800x800 @ 10000 spp =
    6.4 G paths, the headline target).  The oracle cannot render these in test time, so the image is
    box-filtered 5x5 down to the committed low-resolution oracle render -- the mean over a 5x5 pixel
    block of the full-size camera is the same integral as one pixel of the 1/5-size camera -- and
    accepted with the variance-aware bound.  (All strata are needed: a contiguous sub-range of the
    stratified grid covers only a band of each pixel.)"""
    gold = np.load(util.GOLDEN / f"{name}.npz")
    b = BuiltScene(cfg, variant=variant)            # BASELINE defaults: full width / spp / depth
    g = Scene(b)
    H, W = g.info.image_height, g.info.image_width
    gh, gw = gold["mean"].shape[:2]
    assert (H // 5, W // 5) == (gh, gw) and g.info.spp_used in (49, 961, 1936, 10000)
    n = g.info.spp_used
    sg, st = g.render(0, n)
    assert st["paths"] == H * W * n and st["nonfinite_samples"] == 0
    mean = (sg / n)[: gh * 5, : gw * 5].reshape(gh, 5, gw, 5, 3).mean(axis=(1, 3))
    # a 5x5 block of n-sample pixels has (about) the variance of one pixel with 25 n samples
    ok, rep = util.image_acceptance(mean, 25 * n, gold["mean"].astype(np.float64), int(gold["spp"]), gold["var"].astype(np.float64))
    print(name, (W, H), n, rep)
    assert ok, rep
