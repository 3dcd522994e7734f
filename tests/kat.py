"""A third, independent implementation -- plain numpy, written from the reference's formulas (file:line cited at every
function), sharing no code with oracle/oracle.cpp or the CUDA library -- of the closed-form pieces of the path.
tests/test_kat.py holds the oracle (CPU) and the device code (GPU, through the C ABI's rtb_eval_* hooks) to it.
All functions are vectorised over the leading axis."""
import numpy as np


def _dot(a, b):
    return (a * b).sum(axis=-1)


def _unit(a):
    return a / np.linalg.norm(a, axis=-1, keepdims=True)


# ---- Quad (src/object.rs:428-506) ---------------------------------------------------------------
def quad_hit(q, u, v, origin, direction, t_min, t_max=np.inf):
    """Quad::new :428-445 + Quad::hit :453-490 -> (hit mask, t, unit normal)"""
    n = np.cross(u, v)
    normal = n / np.linalg.norm(n)
    w = n / _dot(n, n)
    d = _dot(normal, q)
    denom = _dot(direction, normal)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (d - _dot(origin, normal)) / denom
    p = origin + t[:, None] * direction
    h = p - q
    a = _dot(np.cross(h, v), w)
    b = _dot(np.cross(u, h), w)
    hit = (np.abs(denom) >= 1e-8) & (t >= t_min) & (t <= t_max) & ~((a < 0) | (1 < a) | (b < 0) | (1 < b))
    return hit, t, normal


def quad_pdf_value(q, u, v, origin, direction):
    """Quad::pdf_value :492-501: distance^2 / (|cos| * area) where the probe hits (t_min 0.001), else 0"""
    hit, t, normal = quad_hit(q, u, v, origin, direction, 0.001)
    area = np.linalg.norm(np.cross(u, v))
    dist2 = t * t * _dot(direction, direction)
    cosine = np.abs(_dot(direction, normal) / np.linalg.norm(direction, axis=-1))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(hit, dist2 / (cosine * area), 0.0)


# ---- Sphere (src/object.rs:145-212) ---------------------------------------------------------------
def sphere_hit(center, radius, origin, direction, t_min, t_max=np.inf):
    """Sphere::hit :145-166 (root selection over the OPEN interval, Interval::surrounds)"""
    oc = origin - center
    a = _dot(direction, direction)
    half_b = _dot(oc, direction)
    c = _dot(oc, oc) - radius * radius
    disc = half_b * half_b - a * c
    sq = np.sqrt(np.maximum(disc, 0))
    r1, r2 = (-half_b - sq) / a, (-half_b + sq) / a
    ok1 = (t_min < r1) & (r1 < t_max)
    ok2 = (t_min < r2) & (r2 < t_max)
    t = np.where(ok1, r1, r2)
    return (disc >= 0) & (ok1 | ok2), t


def sphere_pdf_value(center, radius, origin, direction):
    """Sphere::pdf_value :190-202: uniform over the cone the sphere subtends (stationary spheres)"""
    hit, _ = sphere_hit(center, radius, origin, direction, 0.001)
    with np.errstate(invalid="ignore", divide="ignore"):
        cos_theta_max = np.sqrt(1.0 - radius * radius / _dot(center - origin, center - origin))
        return np.where(hit, 1.0 / (2.0 * np.pi * (1.0 - cos_theta_max)), 0.0)


def light_list_pdf(pdfs):
    """HittableList::pdf_value src/hittable.rs:115-124: the arithmetic mean over the listed objects"""
    return np.mean(np.stack(pdfs, axis=0), axis=0)


# ---- Dielectric (src/material.rs:156-191, src/vec3.rs:219-229) ----------------------------------------
def dielectric_direction(direction, normal, front_face, ir, U):
    ratio = np.where(front_face, 1.0 / ir, ir)
    ud = _unit(direction)
    cos_theta = np.minimum(_dot(-ud, normal), 1.0)
    sin_theta = np.sqrt(1.0 - cos_theta * cos_theta)
    cannot = ratio * sin_theta > 1.0
    r0 = ((1 - ratio) / (1 + ratio)) ** 2
    schlick = r0 + (1 - r0) * (1 - cos_theta) ** 5                      # reflectance :156-163
    reflect = ud - 2 * _dot(ud, normal)[:, None] * normal               # vec3.rs:219-221
    perp = ratio[:, None] * (ud + cos_theta[:, None] * normal)          # vec3.rs:223-229
    par = -np.sqrt(np.abs(1.0 - _dot(perp, perp)))[:, None] * normal
    return np.where((cannot | (schlick > U))[:, None], reflect, perp + par)


# ---- textures (src/texture.rs:71-81, 127-130; src/perlin.rs:30-96) -----------------------------------
def checker_is_even(inv_scale, p):
    """CheckerTexture::value :71-81: floor -> i32, Rust's % keeps the sign (-1 % 2 == -1: odd)"""
    s = np.floor(inv_scale * p).astype(np.int64).sum(axis=-1)
    return np.fmod(s, 2) == 0


def perlin_noise(ranvec, perm_x, perm_y, perm_z, p):
    """Perlin::noise :30-54 + trilinear_interp :74-96 (Hermite-smoothed)"""
    fl = np.floor(p)
    u, v, w = (p - fl).T
    i, j, k = fl.astype(np.int64).T
    uu, vv, ww = u * u * (3 - 2 * u), v * v * (3 - 2 * v), w * w * (3 - 2 * w)
    accum = np.zeros(len(p))
    for di in (0, 1):
        for dj in (0, 1):
            for dk in (0, 1):
                c = ranvec[perm_x[(i + di) & 255] ^ perm_y[(j + dj) & 255] ^ perm_z[(k + dk) & 255]]
                weight = np.stack([u - di, v - dj, w - dk], axis=1)
                accum += ((di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww))
                          * _dot(c, weight))
    return accum


def perlin_turb(ranvec, perm_x, perm_y, perm_z, p, depth=7):
    """Perlin::turb_depth :56-68"""
    accum, weight, tp = np.zeros(len(p)), 1.0, p.copy()
    for _ in range(depth):
        accum += weight * perlin_noise(ranvec, perm_x, perm_y, perm_z, tp)
        weight *= 0.5
        tp = tp * 2.0
    return np.abs(accum)


def noise_texture_value(scale, ranvec, perm_x, perm_y, perm_z, p):
    """NoiseTexture::value src/texture.rs:127-130: grey 0.5 (1 + sin(s.z + 10 turb(s)))"""
    s = scale * p
    return 0.5 * (1.0 + np.sin(s[:, 2] + 10.0 * perlin_turb(ranvec, perm_x, perm_y, perm_z, s)))


def image_texture_value(rgb, u, v):
    """ImageTexture::value src/texture.rs:95-107 + RtImage::pixel_data src/rt_image.rs:37-46 (nearest, v flipped, / 255,
    no sRGB decoding).  rgb = (H, W, 3) uint8, top row first."""
    H, W = rgb.shape[:2]
    uc, vc = np.clip(u, 0, 1), np.clip(v, 0, 1)
    i = np.minimum((uc * W).astype(np.int64), W - 1)
    j = (vc * H).astype(np.int64)
    y = np.clip(H - j - 1, 0, H - 1)
    y = np.where(j >= H, H - 1, y)     # `H - j - 1` wraps below zero when j == H (u32), then clamps to H - 1
    return rgb[y, i].astype(np.float64) * (1.0 / 255.0)     # `r * color_scale`, color_scale = 1.0 / 255.0


# ---- colour (src/color.rs:8-59) ---------------------------------------------------------------------
def write_color(pixels, spp, exposure=None):
    x = np.asarray(pixels, dtype=np.float64) / spp
    if exposure is not None:
        x = 1.0 - np.e ** (-exposure * x)
    with np.errstate(invalid="ignore"):
        g = np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(np.maximum(x, 0), 1 / 2.4) - 0.055)
        c = np.where(g < 0, 0.0, np.where(g > 0.999, 0.999, g))
        out = np.where(np.isnan(c), 0.0, 256.0 * c)
    return np.floor(out).astype(np.uint8)


def auto_expose(pixels, spp):
    """auto_expose src/render.rs:325-339 (sequential sum order reproduced with math.fsum-free plain accumulation)"""
    px = np.asarray(pixels, dtype=np.float64).reshape(-1, 3)
    lum = 0.2126 * px[:, 0] + 0.71516 * px[:, 1] + 0.072169 * px[:, 2]
    weight = 1.0 / len(px)
    medium = 0.0
    for term in weight * (lum * lum):
        medium = medium + term
    medium = medium / (spp * spp)
    return -np.log(0.6) / np.sqrt(medium) if medium > 0.001 else 1.0
