"""Thin Python face of librtb200.so (tests/bench plumbing; every call goes through the C ABI)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import RtbRenderParams, RtbSceneInfo, RtbStats


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """rtb_scene* wrapper: upload once, then render / trace / evaluate."""

    def __init__(self, built, device: int = 0):
        self._lib = capi.load_library()
        self._h = C.c_void_p()
        desc = built.desc if hasattr(built, "desc") else built
        capi.check(self._lib, self._lib.rtb_scene_create(desc, device, C.byref(self._h)), "rtb_scene_create")
        self._built = built  # keep the description's buffers alive
        self.info = RtbSceneInfo()
        capi.check(self._lib, self._lib.rtb_scene_info(self._h, C.byref(self.info)), "rtb_scene_info")

    # --- the hot path -------------------------------------------------------------------------
    def render(self, sample_begin: int = 0, sample_end: int | None = None, pipeline: int = 0,
               collect_stats: bool = False, out: np.ndarray | None = None):
        """rtb_render: returns (pixels[h,w,3] f64 sums, stats dict). `out` is accumulated into."""
        h, w = self.info.image_height, self.info.image_width
        if sample_end is None:
            sample_end = self.info.spp_used
        if out is None:
            out = np.zeros((h, w, 3), dtype=np.float64)
        assert out.dtype == np.float64 and out.shape == (h, w, 3) and out.flags.c_contiguous
        params = RtbRenderParams(sample_begin, sample_end, pipeline, 1 if collect_stats else 0)
        stats = RtbStats()
        capi.check(self._lib, self._lib.rtb_render(self._h, C.byref(params), _ptr(out), C.byref(stats)), "rtb_render")
        return out, stats.as_dict()

    def set_option(self, option: int, value: int):
        """rtb_scene_set_option (capi.OPT_*): per-scene tuning knobs / A-B arms"""
        capi.check(self._lib, self._lib.rtb_scene_set_option(self._h, option, int(value)), "rtb_scene_set_option")
        return self

    def accum_to_pixels(self, d_accum_ptr: int, out: np.ndarray | None = None) -> np.ndarray:
        """rtb_accum_to_pixels: a device accumulation buffer (fixed-point u64 x 4) -> += host f64 sums"""
        h, w = self.info.image_height, self.info.image_width
        if out is None:
            out = np.zeros((h, w, 3), dtype=np.float64)
        capi.check(self._lib, self._lib.rtb_accum_to_pixels(self._h, C.c_void_p(d_accum_ptr), _ptr(out)), "rtb_accum_to_pixels")
        return out

    def eval_dielectric(self, in9: np.ndarray) -> np.ndarray:
        """Dielectric::scatter on the device: n x {d[3], face normal[3], front, ir, u} -> n x direction[3]"""
        a = np.ascontiguousarray(in9, dtype=np.float64).reshape(-1, 9)
        out = np.zeros((len(a), 3))
        capi.check(self._lib, self._lib.rtb_eval_dielectric(self._h, _ptr(a), len(a), _ptr(out)), "rtb_eval_dielectric")
        return out

    def checkpoint_save(self, d_accum_ptr: int, path: str):
        capi.check(self._lib, self._lib.rtb_checkpoint_save(self._h, C.c_void_p(d_accum_ptr), str(path).encode()), "rtb_checkpoint_save")

    def checkpoint_load(self, d_accum_ptr: int, path: str):
        capi.check(self._lib, self._lib.rtb_checkpoint_load(self._h, C.c_void_p(d_accum_ptr), str(path).encode()), "rtb_checkpoint_load")

    def philox(self, ctr_key: np.ndarray) -> np.ndarray:
        """the device's Philox4x32-10 on n x {c0,c1,c2,c3,k0,k1} (Random123 known-answer hook)"""
        ck = np.ascontiguousarray(ctr_key, dtype=np.uint32).reshape(-1, 6)
        out = np.zeros((len(ck), 4), dtype=np.uint32)
        capi.check(self._lib, self._lib.rtb_philox(self._h, _ptr(ck), len(ck), _ptr(out)), "rtb_philox")
        return out

    def render_device(self, d_accum_ptr: int, sample_begin: int, sample_end: int, stream: int = 0,
                      pipeline: int = 0, collect_stats: bool = False):
        """rtb_render_device: accumulate into a caller-owned device buffer of w*h x 4 uint64 (fixed-point sums), asynchronously."""
        params = RtbRenderParams(sample_begin, sample_end, pipeline, 1 if collect_stats else 0)
        capi.check(self._lib, self._lib.rtb_render_device(self._h, C.byref(params), C.c_void_p(d_accum_ptr),
                                                          C.c_void_p(stream)), "rtb_render_device")

    def render_stats(self) -> dict:
        stats = RtbStats()
        capi.check(self._lib, self._lib.rtb_render_stats(self._h, C.byref(stats)), "rtb_render_stats")
        return stats.as_dict()

    # --- parity harness -----------------------------------------------------------------------
    def camera_rays(self) -> np.ndarray:
        rays = np.zeros(self.info.image_height * self.info.image_width, dtype=capi.RAY_DTYPE)
        capi.check(self._lib, self._lib.rtb_camera_rays(self._h, _ptr(rays)), "rtb_camera_rays")
        return rays

    def trace(self, rays: np.ndarray, flags: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=capi.HIT_DTYPE)
        capi.check(self._lib, self._lib.rtb_trace(self._h, _ptr(rays), len(rays), flags, _ptr(hits)), "rtb_trace")
        return hits

    def medium_interval(self, medium: int, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        t0 = np.zeros(len(rays)); t1 = np.zeros(len(rays))
        capi.check(self._lib, self._lib.rtb_medium_interval(self._h, medium, _ptr(rays), len(rays), _ptr(t0), _ptr(t1)),
                   "rtb_medium_interval")
        return t0, t1

    def eval_texture(self, texture: int, uvp: np.ndarray) -> np.ndarray:
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((len(uvp), 3))
        capi.check(self._lib, self._lib.rtb_eval_texture(self._h, texture, _ptr(uvp), len(uvp), _ptr(out)), "rtb_eval_texture")
        return out

    def eval_light_pdf(self, origin_dir: np.ndarray) -> np.ndarray:
        od = np.ascontiguousarray(origin_dir, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(od))
        capi.check(self._lib, self._lib.rtb_eval_light_pdf(self._h, _ptr(od), len(od), _ptr(out)), "rtb_eval_light_pdf")
        return out

    def write_color(self, pixels: np.ndarray, spp: float, exposure: float = 0.0) -> np.ndarray:
        px = np.ascontiguousarray(pixels, dtype=np.float64)
        out = np.zeros(px.shape, dtype=np.uint8)
        capi.check(self._lib, self._lib.rtb_write_color(self._h, _ptr(px), px.size // 3, float(spp), float(exposure), _ptr(out)),
                   "rtb_write_color")
        return out

    def close(self):
        if self._h:
            self._lib.rtb_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_multi(built, n_devices: int, sample_begin: int = 0, sample_end: int | None = None, pipeline: int = 0,
                 devices=None, collect_stats: bool = False, out: np.ndarray | None = None):
    """rtb_render_multi: the reference seam on the GPUs of one box (one thread + stream per GPU, one NCCL sum-reduce of
    the int64 accumulation buffers, one device-to-host copy).  Returns (pixels[h,w,3] f64 sums, stats dict)."""
    lib = capi.load_library()
    desc = built.desc if hasattr(built, "desc") else built
    cam = desc.contents.camera
    h = max(1, int(cam.image_width / cam.aspect_ratio))
    root = int(np.sqrt(cam.samples_per_pixel))
    if sample_end is None:
        sample_end = root * root
    if out is None:
        out = np.zeros((h, cam.image_width, 3), dtype=np.float64)
    params = RtbRenderParams(sample_begin, sample_end, pipeline, 1 if collect_stats else 0)
    stats = RtbStats()
    devs = None
    if devices is not None:
        devs = (C.c_int * len(devices))(*devices)
    capi.check(lib, lib.rtb_render_multi(desc, n_devices, devs, C.byref(params), _ptr(out), C.byref(stats)), "rtb_render_multi")
    return out, stats.as_dict()


def auto_expose(pixels: np.ndarray, spp: float) -> float:
    lib = capi.load_library()
    px = np.ascontiguousarray(pixels, dtype=np.float64)
    e = C.c_double()
    capi.check(lib, lib.rtb_auto_expose(_ptr(px), px.size // 3, float(spp), C.byref(e)), "rtb_auto_expose")
    return e.value
