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


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4", "c5"])
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


def test_medium_intervals_and_textures_and_light_pdf():
    b = BuiltScene("c4", width=64, spp=4, variant=1)
    o, e = orc.OracleScene(b), EmuScene(b)
    rays = o.camera_rays()
    for m in range(2):
        a0, a1 = o.medium_interval(m, rays)
        b0, b1 = e.medium_interval(m, rays)
        assert np.array_equal(np.isnan(a0), np.isnan(b0))
        ok = ~np.isnan(a0)
        assert np.allclose(a0[ok], b0[ok], rtol=1e-10) and np.allclose(a1[ok], b1[ok], rtol=1e-10)
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
