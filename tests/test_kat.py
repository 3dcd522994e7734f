"""Known answers from a THIRD implementation (tests/kat.py: numpy, written from the reference's formulas) for the
closed-form pieces of the path: the oracle is checked against it on the CPU, the device code on the GPU.  Together with
the Random123 vectors and the reference's own images this widens what pins the oracle beyond its author's reading."""
import ctypes as C

import numpy as np
import pytest

from oracle import orc
from surely_raytracing_b200 import capi
from surely_raytracing_b200.scenes import BuiltScene
from tests import kat

CORNELL_LIGHT = (np.array([343., 554., 332.]), np.array([-130., 0., 0.]), np.array([0., 0., -105.]))   # src/main.rs:441-446
CORNELL_SPHERE = (np.array([190., 90., 190.]), 90.0)                                                     # src/main.rs:482-483


def _probes(rng, n):
    origin = rng.uniform(30, 520, (n, 3))
    toward = np.where(rng.random((n, 1)) < 0.5, CORNELL_LIGHT[0] + rng.random((n, 1)) * CORNELL_LIGHT[1] + rng.random((n, 1)) * CORNELL_LIGHT[2],
                      CORNELL_SPHERE[0] + rng.normal(size=(n, 3)) * 60)
    direction = (toward - origin) * rng.uniform(0.2, 3.0, (n, 1))
    direction[: n // 8] = rng.normal(size=(n // 8, 3))
    return origin, direction


def _dielectric_cases(rng, n):
    d = rng.normal(size=(n, 3)) * rng.uniform(0.1, 5, (n, 1))
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where((np.einsum("ij,ij->i", d, nrm) > 0)[:, None], -nrm, nrm)           # face normal: against the ray
    graze = rng.random(n) < 0.3                                                         # plenty of near-grazing incidence: total internal reflection
    t = np.cross(nrm, rng.normal(size=(n, 3)))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    d = np.where(graze[:, None], t - nrm * rng.uniform(0.01, 0.6, (n, 1)), d)
    front = rng.random(n) < 0.5
    ir = rng.choice([1.5, 1.33, 2.4, 1.0], n)
    U = rng.random(n)
    return np.column_stack([d, nrm, front.astype(float), ir, U]), d, nrm, front, ir, U


def _perlin_scene():
    """c4's noise texture (index by kind) and its tables, as the description holds them"""
    b = BuiltScene("c4", width=32, spp=4)
    d = b.desc.contents
    ti = next(i for i in range(d.n_textures) if d.textures[i].kind == 3)
    p = d.perlins[d.textures[ti].a]
    ranvec = np.array([[p.ranvec[i][a] for a in range(3)] for i in range(256)])
    perms = [np.array(list(getattr(p, f"perm_{ax}")), dtype=np.int64) for ax in "xyz"]
    return b, ti, d.textures[ti].scale, ranvec, perms


def test_oracle_light_pdfs_match_the_numpy_formulas():
    rng = np.random.default_rng(2)
    origin, direction = _probes(rng, 40000)
    want_q = kat.quad_pdf_value(*CORNELL_LIGHT, origin, direction)
    want_s = kat.sphere_pdf_value(*CORNELL_SPHERE, origin, direction)
    assert (want_q > 0).sum() > 5000 and (want_s > 0).sum() > 5000
    od = np.hstack([origin, direction])
    got5 = orc.OracleScene(BuiltScene("c5", width=32, spp=4)).eval_light_pdf(od)      # lights = [quad, sphere]
    got2 = orc.OracleScene(BuiltScene("c2", width=32, spp=4)).eval_light_pdf(od)      # lights = [quad]
    want5 = kat.light_list_pdf([want_q, want_s])
    fin = np.isfinite(want5)
    assert np.array_equal(np.isfinite(got5), fin)
    assert np.allclose(got5[fin], want5[fin], rtol=1e-12, atol=0) and np.allclose(got2, want_q, rtol=1e-12, atol=0)


def test_oracle_dielectric_matches_the_numpy_formulas():
    rng = np.random.default_rng(3)
    in9, d, nrm, front, ir, U = _dielectric_cases(rng, 50000)
    want = kat.dielectric_direction(d, nrm, front, ir, U)
    got = orc.eval_dielectric(in9)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-13)
    refl = np.abs(np.einsum("ij,ij->i", want, nrm) - np.einsum("ij,ij->i", -d / np.linalg.norm(d, axis=1, keepdims=True), nrm)) < 1e-9
    assert 0.2 < refl.mean() < 0.9                                                      # both branches well populated


def test_oracle_textures_match_the_numpy_formulas():
    rng = np.random.default_rng(4)
    b, ti, scale, ranvec, perms = _perlin_scene()
    o = orc.OracleScene(b)
    p = rng.uniform(-400, 700, (20000, 3))
    p[:2000] = np.round(p[:2000])                                                       # lattice points: u = v = w = 0
    uvp = np.hstack([rng.random((len(p), 2)), p])
    got = o.eval_texture(ti, uvp)
    want = kat.noise_texture_value(scale, ranvec, *perms, p)
    assert np.allclose(got, want[:, None].repeat(3, 1), rtol=1e-11, atol=1e-12)
    # checker: c1's ground (inv_scale 1/0.32), negative coordinates included (Rust's % keeps the sign)
    b1 = BuiltScene("c1", width=32, spp=4)
    d1 = b1.desc.contents
    ci = next(i for i in range(d1.n_textures) if d1.textures[i].kind == 1)
    tex = d1.textures[ci]
    even = np.array(list(d1.textures[tex.a].color)), np.array(list(d1.textures[tex.b].color))
    pc = rng.uniform(-12, 12, (20000, 3))
    got = orc.OracleScene(b1).eval_texture(ci, np.hstack([rng.random((len(pc), 2)), pc]))
    want = np.where(kat.checker_is_even(tex.scale, pc)[:, None], even[0], even[1])
    assert np.array_equal(got, want)
    # image: c4's synthetic earth, u, v beyond [0, 1] included
    d = b.desc.contents
    ii = next(i for i in range(d.n_textures) if d.textures[i].kind == 2)
    im = d.images[d.textures[ii].a]
    rgb = np.ctypeslib.as_array(im.rgb, shape=(im.height, im.width, 3))
    uv = rng.uniform(-0.2, 1.2, (20000, 2))
    uv[:200] = rng.choice([0.0, 1.0], (200, 2))
    got = o.eval_texture(ii, np.hstack([uv, p]))
    assert np.array_equal(got, kat.image_texture_value(rgb, uv[:, 0], uv[:, 1]))


def test_oracle_output_stage_matches_the_numpy_formulas():
    rng = np.random.default_rng(5)
    px = np.abs(rng.normal(size=(20000, 3))) * rng.choice([0.001, 0.1, 1.0, 30.0], size=(20000, 1)) * 9
    px[7] = np.nan
    assert np.array_equal(orc.write_color(px, 9.0), kat.write_color(px, 9.0))
    assert np.array_equal(orc.write_color(px, 9.0, 0.7), kat.write_color(px, 9.0, 0.7))
    px[7] = 0
    assert orc.auto_expose(px, 9.0) == kat.auto_expose(px, 9.0)
    assert orc.auto_expose(px * 1e-3, 9.0) == 1.0 == kat.auto_expose(px * 1e-3, 9.0)


# ---- the device code against the same answers -------------------------------------------------------------------
@pytest.mark.gpu
def test_device_light_pdfs_dielectric_and_textures_match_the_numpy_formulas():
    from surely_raytracing_b200 import Scene
    rng = np.random.default_rng(6)
    origin, direction = _probes(rng, 40000)
    od = np.hstack([origin, direction])
    want5 = kat.light_list_pdf([kat.quad_pdf_value(*CORNELL_LIGHT, origin, direction), kat.sphere_pdf_value(*CORNELL_SPHERE, origin, direction)])
    got5 = Scene(BuiltScene("c5", width=32, spp=4)).eval_light_pdf(od)
    fin = np.isfinite(want5)
    assert np.array_equal(np.isfinite(got5), fin) and np.allclose(got5[fin], want5[fin], rtol=1e-9, atol=0)
    # Dielectric::scatter: fp32 on the device.  Away from the two decision boundaries (total internal reflection,
    # Schlick vs the draw) the branch must be the same and the direction equal to fp32 accuracy.
    in9, d, nrm, front, ir, U = _dielectric_cases(rng, 50000)
    g = Scene(BuiltScene("c2", width=16, spp=4))
    got = g.eval_dielectric(in9)
    want = kat.dielectric_direction(d, nrm, front, ir, U)
    ratio = np.where(front, 1 / ir, ir)
    ud = d / np.linalg.norm(d, axis=1, keepdims=True)
    cos_t = np.minimum(np.einsum("ij,ij->i", -ud, nrm), 1.0)
    sin_t = np.sqrt(1 - cos_t ** 2)
    r0 = ((1 - ratio) / (1 + ratio)) ** 2
    margin = np.minimum(np.abs(ratio * sin_t - 1.0), np.abs(r0 + (1 - r0) * (1 - cos_t) ** 5 - U))
    clear = margin > 1e-4
    # |1 - |perp|^2| under the square root amplifies fp32 rounding near the critical angle: compare where it is not tiny
    perp2 = (ratio ** 2) * (1 - cos_t ** 2)
    clear &= np.abs(1 - perp2) > 1e-3
    assert clear.mean() > 0.95
    assert np.abs(got[clear] - want[clear]).max() < 5e-5
    # textures: Perlin / image in fp32 on the device (2e-4), checker exact
    b, ti, scale, ranvec, perms = _perlin_scene()
    g4 = Scene(b)
    p = rng.uniform(-400, 700, (20000, 3))
    uv = rng.uniform(-0.2, 1.2, (20000, 2))
    got = g4.eval_texture(ti, np.hstack([uv, p]))
    want = kat.noise_texture_value(scale, ranvec, *perms, p)
    assert np.abs(got - want[:, None]).max() < 2e-4
    dsc = b.desc.contents
    ii = next(i for i in range(dsc.n_textures) if dsc.textures[i].kind == 2)
    im = dsc.images[dsc.textures[ii].a]
    rgb = np.ctypeslib.as_array(im.rgb, shape=(im.height, im.width, 3))
    got = g4.eval_texture(ii, np.hstack([uv, p]))
    want = kat.image_texture_value(rgb, uv[:, 0].astype(np.float32).astype(np.float64), uv[:, 1].astype(np.float32).astype(np.float64))
    same = np.abs(got - want).max(axis=1) < 1e-6
    assert same.mean() > 0.999                                                          # (a texel boundary hit within fp32 rounding of u * W)
    px = np.abs(rng.normal(size=(20000, 3))) * rng.choice([0.001, 0.1, 1.0, 30.0], size=(20000, 1)) * 9
    assert np.array_equal(g.write_color(px, 9.0, 0.7), kat.write_color(px, 9.0, 0.7))
