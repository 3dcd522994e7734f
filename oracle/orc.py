"""ctypes loader of oracle/liboracle.so -- TEST INFRASTRUCTURE (see oracle/oracle.cpp).
Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from surely_raytracing_b200 import capi

ORACLE_DIR = Path(__file__).resolve().parent
LIB_PATH = ORACLE_DIR / "liboracle.so"
SAMPLER_REF, SAMPLER_KEYED = 0, 1


class OrcStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("node_visits", C.c_uint64),
                ("prim_tests", C.c_uint64), ("medium_probes", C.c_uint64), ("nonfinite_samples", C.c_uint64),
                ("wall_ms", C.c_double), ("threads", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


def build(force: bool = False):
    src = ORACLE_DIR / "oracle.cpp"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B", "liboracle.so"], check=True, capture_output=True)


_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        lib = C.CDLL(str(LIB_PATH))
        vp, i32, i64, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
        lib.orc_last_error.restype = C.c_char_p
        lib.orc_scene_create.argtypes = [C.POINTER(capi.RtbSceneDesc), C.POINTER(vp)]
        lib.orc_scene_destroy.argtypes = [vp]; lib.orc_scene_destroy.restype = None
        lib.orc_set_use_bvh.argtypes = [vp, C.c_int]; lib.orc_set_use_bvh.restype = None
        lib.orc_scene_info.argtypes = [vp, C.POINTER(capi.RtbSceneInfo)]
        lib.orc_render.argtypes = [vp, i64, i64, C.c_int, C.c_int, vp, vp, C.POINTER(OrcStats)]
        lib.orc_trace.argtypes = [vp, vp, i64, u32, vp]
        lib.orc_camera_rays.argtypes = [vp, vp]
        lib.orc_medium_interval.argtypes = [vp, i32, vp, i64, vp, vp]
        lib.orc_eval_texture.argtypes = [vp, i32, vp, i64, vp]
        lib.orc_eval_light_pdf.argtypes = [vp, vp, i64, vp]
        lib.orc_sample_lights.argtypes = [vp, vp, i64, C.c_int, u64, vp]
        lib.orc_sample_directions.argtypes = [C.c_int, i64, C.c_int, u64, vp]
        lib.orc_sphere_uv.argtypes = [vp, vp]; lib.orc_sphere_uv.restype = None
        lib.orc_philox.argtypes = [vp, vp, vp]; lib.orc_philox.restype = None
        lib.orc_write_color.argtypes = [vp, i64, C.c_double, C.c_double, vp]
        lib.orc_eval_dielectric.argtypes = [vp, i64, vp]
        lib.orc_auto_expose.argtypes = [vp, i64, C.c_double]; lib.orc_auto_expose.restype = C.c_double
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(lib, rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib.orc_last_error().decode()}")


class OracleScene:
    def __init__(self, built, use_bvh: bool = True):
        self._lib = load()
        self._h = C.c_void_p()
        desc = built.desc if hasattr(built, "desc") else built
        _check(self._lib, self._lib.orc_scene_create(desc, C.byref(self._h)), "orc_scene_create")
        self._built = built
        self._lib.orc_set_use_bvh(self._h, 1 if use_bvh else 0)
        self.info = capi.RtbSceneInfo()
        self._lib.orc_scene_info(self._h, C.byref(self.info))

    def set_use_bvh(self, on: bool):
        self._lib.orc_set_use_bvh(self._h, 1 if on else 0)

    def render(self, sample_begin=0, sample_end=None, sampler=SAMPLER_REF, threads=0, want_sumsq=False):
        h, w = self.info.image_height, self.info.image_width
        if sample_end is None:
            sample_end = self.info.spp_used
        s = np.zeros((h, w, 3)); s2 = np.zeros((h, w, 3)) if want_sumsq else None
        st = OrcStats()
        _check(self._lib, self._lib.orc_render(self._h, sample_begin, sample_end, sampler, threads, _ptr(s),
                                               _ptr(s2) if want_sumsq else None, C.byref(st)), "orc_render")
        return (s, s2, st.as_dict()) if want_sumsq else (s, st.as_dict())

    def camera_rays(self):
        rays = np.zeros(self.info.image_height * self.info.image_width, dtype=capi.RAY_DTYPE)
        _check(self._lib, self._lib.orc_camera_rays(self._h, _ptr(rays)), "orc_camera_rays")
        return rays

    def trace(self, rays, flags=0):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=capi.HIT_DTYPE)
        _check(self._lib, self._lib.orc_trace(self._h, _ptr(rays), len(rays), flags, _ptr(hits)), "orc_trace")
        return hits

    def medium_interval(self, medium, rays):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        t0 = np.zeros(len(rays)); t1 = np.zeros(len(rays))
        _check(self._lib, self._lib.orc_medium_interval(self._h, medium, _ptr(rays), len(rays), _ptr(t0), _ptr(t1)),
               "orc_medium_interval")
        return t0, t1

    def eval_texture(self, texture, uvp):
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((len(uvp), 3))
        _check(self._lib, self._lib.orc_eval_texture(self._h, texture, _ptr(uvp), len(uvp), _ptr(out)), "orc_eval_texture")
        return out

    def eval_light_pdf(self, origin_dir):
        od = np.ascontiguousarray(origin_dir, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(od))
        _check(self._lib, self._lib.orc_eval_light_pdf(self._h, _ptr(od), len(od), _ptr(out)), "orc_eval_light_pdf")
        return out

    def sample_lights(self, origin, n, sampler=SAMPLER_REF, stream=1):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        out = np.zeros((n, 3))
        _check(self._lib, self._lib.orc_sample_lights(self._h, _ptr(o), n, sampler, stream, _ptr(out)), "orc_sample_lights")
        return out

    def close(self):
        if self._h:
            self._lib.orc_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample_directions(which: int, n: int, sampler=SAMPLER_REF, seed=1):
    out = np.zeros((n, 3))
    load().orc_sample_directions(which, n, sampler, seed, _ptr(out))
    return out


def sphere_uv(p):
    p = np.ascontiguousarray(p, dtype=np.float64); uv = np.zeros(2)
    load().orc_sphere_uv(_ptr(p), _ptr(uv))
    return uv


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
    load().orc_philox(_ptr(c), _ptr(k), _ptr(o))
    return o


def write_color(pixels, spp, exposure=0.0):
    px = np.ascontiguousarray(pixels, dtype=np.float64)
    out = np.zeros(px.shape, dtype=np.uint8)
    load().orc_write_color(_ptr(px), px.size // 3, float(spp), float(exposure), _ptr(out))
    return out


def auto_expose(pixels, spp):
    """auto_expose (reference src/render.rs:325-339) over f64 pixel sums"""
    px = np.ascontiguousarray(pixels, dtype=np.float64)
    return float(load().orc_auto_expose(_ptr(px), px.size // 3, float(spp)))


def eval_dielectric(in9):
    a = np.ascontiguousarray(in9, dtype=np.float64).reshape(-1, 9)
    out = np.zeros((len(a), 3))
    load().orc_eval_dielectric(_ptr(a), len(a), _ptr(out))
    return out
