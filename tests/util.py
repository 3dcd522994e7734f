"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np

GOLDEN = __import__("pathlib").Path(__file__).resolve().parent / "golden"

CONFIG_VARIANTS = [("c1", 0), ("c2", 0), ("c3", 0), ("c3", 1), ("c4", 0), ("c4", 1), ("c5", 0)]


def hit_errors(ref, got):
    """first-hit comparison of two HIT_DTYPE arrays: id mismatches and max relative errors."""
    mism = int((ref["prim"] != got["prim"]).sum())
    both = (ref["prim"] >= 0) & (ref["prim"] == got["prim"])
    if not both.any():
        return mism, 0.0, 0.0, 0.0
    t_rel = np.abs(ref["t"][both] - got["t"][both]) / np.abs(ref["t"][both])
    dn = np.linalg.norm(ref["normal"][both] - got["normal"][both], axis=1)
    duv = np.maximum(np.abs(ref["u"][both] - got["u"][both]), np.abs(ref["v"][both] - got["v"][both]))
    return mism, float(t_rel.max()), float(dn.max()), float(duv.max())


def secondary_rays(hits, rng, n_max=20000, t_min=1e-4):
    """Seeded random secondary rays leaving the surfaces found by `hits` (exercises t_min and the
    inside-sphere roots): origin = hit point, direction = random unit vector (both hemispheres)."""
    from surely_raytracing_b200 import capi
    ok = np.flatnonzero(hits["prim"] >= 0)
    if len(ok) > n_max:
        ok = rng.choice(ok, n_max, replace=False)
    d = rng.normal(size=(len(ok), 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.uniform(0.5, 3.0, size=(len(ok), 1))
    rays = np.zeros(len(ok), dtype=capi.RAY_DTYPE)
    rays["origin"] = hits["p"][ok]
    rays["direction"] = d
    rays["time"] = rng.uniform(0, 1, size=len(ok))
    rays["t_min"] = t_min
    return rays


def image_acceptance(mean_a, n_a, mean_b, n_b, var_px, clamp=10.0):
    """Variance-aware acceptance of SURVEY 8(d): per channel
         RMSE(a-b) <= 1.15 * sqrt(mean_px var * (1/n_a + 1/n_b))      and
         |mean_px(a-b)| <= 4 * sqrt(mean_px var * (1/n_a + 1/n_b) / n_px)
    in linear radiance clamped to [0, clamp].  Returns (ok, report dict)."""
    a = np.clip(mean_a, 0, clamp)
    b = np.clip(mean_b, 0, clamp)
    n_px = a.shape[0] * a.shape[1]
    rep = {}
    ok = True
    for c in range(3):
        sigma2 = float(var_px[..., c].mean()) * (1.0 / n_a + 1.0 / n_b)
        rmse = float(np.sqrt(((a[..., c] - b[..., c]) ** 2).mean()))
        bias = float((a[..., c] - b[..., c]).mean())
        lim_rmse = 1.15 * np.sqrt(sigma2)
        lim_bias = 4.0 * np.sqrt(sigma2 / n_px)
        rep[c] = dict(rmse=rmse, lim_rmse=float(lim_rmse), bias=bias, lim_bias=float(lim_bias))
        ok &= rmse <= lim_rmse and abs(bias) <= lim_bias
    return ok, rep


def psnr8(a8, b8):
    mse = ((a8.astype(np.float64) - b8.astype(np.float64)) ** 2).mean()
    return float(10 * np.log10(255.0 ** 2 / max(mse, 1e-12)))
