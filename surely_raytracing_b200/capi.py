"""ctypes mirror of include/rtb200.h and the loader of librtb200.so.

This is plumbing for tests/ and bench.py: the product is the shared library.  There is no CPU
fallback -- if the CUDA library is missing or no sm_100 device is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("RTB200_LIB", PKG_DIR / "librtb200.so"))  # override: debug builds only
SCENES_LIB_PATH = PKG_DIR / "librtb_scenes.so"

RTB_ABI_VERSION = 2
RTB_FLAG_ISO_PDF_ZERO = 1
RTB_FLAG_PROPAGATE_NAN = 2
RTB_FLAG_BVH4, RTB_FLAG_QNODES, RTB_FLAG_BVH_LEAF4, RTB_FLAG_NO_BOX_SCAN, RTB_FLAG_SUN_LIGHT = 0x10, 0x20, 0x40, 0x80, 0x100
RTB_FLAG_RUSSIAN_ROULETTE, RTB_FLAG_NO_BOX_LEAVES = 0x200, 0x400
RTB_TRACE_BRUTE_FORCE, RTB_TRACE_WAVEFRONT, RTB_TRACE_SECONDARY = 1, 2, 4
(OPT_WF_CAPACITY, OPT_EXACT_LEAVES, OPT_SMEM_TOP, OPT_NO_DEFER_RARE, OPT_EXTEND_BLOCKS, OPT_FINISH_BELOW, OPT_PROFILE,
 OPT_MEGA_BELOW) = range(1, 9)
ACCUM_SCALE = 4294967296.0
PIPELINE_DEFAULT, PIPELINE_MEGAKERNEL, PIPELINE_WAVEFRONT = 0, 1, 2
VARIANT_LIGHTS = 1


class RtbObject(C.Structure):
    _fields_ = [("kind", C.c_int32), ("material", C.c_int32), ("first", C.c_int32), ("count", C.c_int32),
                ("v", C.c_double * 10)]


class RtbMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("texture", C.c_int32), ("color", C.c_double * 3), ("param", C.c_double)]


class RtbTexture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("reserved", C.c_int32),
                ("color", C.c_double * 3), ("scale", C.c_double)]


class RtbImage(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_uint8))]


class RtbPerlin(C.Structure):
    _fields_ = [("ranvec", (C.c_double * 3) * 256), ("perm_x", C.c_int32 * 256), ("perm_y", C.c_int32 * 256),
                ("perm_z", C.c_int32 * 256)]


class RtbCamera(C.Structure):
    _fields_ = [("aspect_ratio", C.c_double), ("image_width", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("max_depth", C.c_int32), ("reserved", C.c_int32), ("vfov", C.c_double),
                ("lookfrom", C.c_double * 3), ("lookat", C.c_double * 3), ("vup", C.c_double * 3),
                ("defocus_angle", C.c_double), ("focus_dist", C.c_double), ("background", C.c_double * 3)]


class RtbSun(C.Structure):
    _fields_ = [("direction", C.c_double * 3), ("albedo", C.c_double * 3), ("angular_diameter", C.c_double)]


class RtbSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("flags", C.c_uint32), ("seed", C.c_uint64),
                ("objects", C.POINTER(RtbObject)), ("n_objects", C.c_int32),
                ("children", C.POINTER(C.c_int32)), ("n_children", C.c_int32), ("world", C.c_int32),
                ("lights", C.POINTER(C.c_int32)), ("n_lights", C.c_int32),
                ("materials", C.POINTER(RtbMaterial)), ("n_materials", C.c_int32),
                ("textures", C.POINTER(RtbTexture)), ("n_textures", C.c_int32),
                ("images", C.POINTER(RtbImage)), ("n_images", C.c_int32),
                ("perlins", C.POINTER(RtbPerlin)), ("n_perlins", C.c_int32),
                ("camera", RtbCamera),
                ("suns", C.POINTER(RtbSun)), ("n_suns", C.c_int32), ("reserved", C.c_int32)]


class RtbRay(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("direction", C.c_double * 3), ("time", C.c_double),
                ("t_min", C.c_double)]


class RtbHit(C.Structure):
    _fields_ = [("prim", C.c_int32), ("front_face", C.c_int32), ("material", C.c_int32), ("reserved", C.c_int32),
                ("t", C.c_double), ("p", C.c_double * 3), ("normal", C.c_double * 3), ("u", C.c_double),
                ("v", C.c_double)]


class RtbSceneInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("image_width", "image_height", "spp_used", "sqrt_spp", "max_depth", "n_surface_prims",
                 "n_boundary_prims", "n_media", "n_bvh_nodes", "n_lights", "bvh_depth", "device")]


class RtbRenderParams(C.Structure):
    _fields_ = [("sample_begin", C.c_int64), ("sample_end", C.c_int64), ("pipeline", C.c_int32),
                ("collect_stats", C.c_int32)]


class RtbStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("node_visits", C.c_uint64),
                ("prim_tests", C.c_uint64), ("medium_probes", C.c_uint64), ("nonfinite_samples", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_ms", C.c_double), ("exact_tests", C.c_uint64),
                ("overflow_rays", C.c_uint64), ("stage_ms", C.c_double * 3)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["stage_ms"] = list(self.stage_ms)
        return d


# numpy views of the ray / hit records (same layout as the C structs: all 8-byte aligned)
RAY_DTYPE = np.dtype([("origin", "f8", 3), ("direction", "f8", 3), ("time", "f8"), ("t_min", "f8")])
HIT_DTYPE = np.dtype([("prim", "i4"), ("front_face", "i4"), ("material", "i4"), ("reserved", "i4"),
                      ("t", "f8"), ("p", "f8", 3), ("normal", "f8", 3), ("u", "f8"), ("v", "f8")])
assert RAY_DTYPE.itemsize == C.sizeof(RtbRay) and HIT_DTYPE.itemsize == C.sizeof(RtbHit)

# every symbol include/rtb200.h declares (tests check that the library exports all of them)
EXPORTS = ["rtb_version", "rtb_device_count", "rtb_last_error", "rtb_scene_create", "rtb_scene_destroy",
           "rtb_scene_info", "rtb_render", "rtb_render_device", "rtb_render_stats", "rtb_trace",
           "rtb_camera_rays", "rtb_medium_interval", "rtb_eval_texture", "rtb_eval_light_pdf",
           "rtb_write_color", "rtb_accum_to_pixels", "rtb_render_multi", "rtb_scene_set_option", "rtb_trim_cache",
           "rtb_auto_expose", "rtb_philox", "rtb_eval_dielectric", "rtb_checkpoint_save", "rtb_checkpoint_load"]


class RtbError(RuntimeError):
    pass


_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """dlopen librtb200.so (built in-tree by __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RtbError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    lib.rtb_version.restype = C.c_int
    lib.rtb_device_count.restype = C.c_int
    lib.rtb_last_error.restype = C.c_char_p
    lib.rtb_scene_create.argtypes = [C.POINTER(RtbSceneDesc), C.c_int, C.POINTER(vp)]
    lib.rtb_scene_destroy.argtypes = [vp]
    lib.rtb_scene_destroy.restype = None
    lib.rtb_scene_info.argtypes = [vp, C.POINTER(RtbSceneInfo)]
    lib.rtb_render.argtypes = [vp, C.POINTER(RtbRenderParams), vp, C.POINTER(RtbStats)]
    lib.rtb_render_device.argtypes = [vp, C.POINTER(RtbRenderParams), vp, vp]
    lib.rtb_render_stats.argtypes = [vp, C.POINTER(RtbStats)]
    lib.rtb_trace.argtypes = [vp, vp, i64, u32, vp]
    lib.rtb_camera_rays.argtypes = [vp, vp]
    lib.rtb_medium_interval.argtypes = [vp, i32, vp, i64, vp, vp]
    lib.rtb_eval_texture.argtypes = [vp, i32, vp, i64, vp]
    lib.rtb_eval_light_pdf.argtypes = [vp, vp, i64, vp]
    lib.rtb_write_color.argtypes = [vp, vp, i64, C.c_double, C.c_double, vp]
    lib.rtb_accum_to_pixels.argtypes = [vp, vp, vp]
    lib.rtb_render_multi.argtypes = [C.POINTER(RtbSceneDesc), C.c_int, vp, C.POINTER(RtbRenderParams), vp, C.POINTER(RtbStats)]
    lib.rtb_scene_set_option.argtypes = [vp, C.c_int, i64]
    lib.rtb_trim_cache.restype = i64
    lib.rtb_auto_expose.argtypes = [vp, i64, C.c_double, C.POINTER(C.c_double)]
    lib.rtb_philox.argtypes = [vp, vp, i64, vp]
    lib.rtb_eval_dielectric.argtypes = [vp, vp, i64, vp]
    lib.rtb_checkpoint_save.argtypes = [vp, vp, C.c_char_p]
    lib.rtb_checkpoint_load.argtypes = [vp, vp, C.c_char_p]
    if path is None:
        _lib = lib
    return lib


def check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.rtb_last_error()
        raise RtbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


_scenes_lib = None


def load_scenes_library() -> C.CDLL:
    global _scenes_lib
    if _scenes_lib is None:
        if not SCENES_LIB_PATH.exists():
            raise RtbError(f"{SCENES_LIB_PATH} is missing: run __graft_entry__.build()")
        lib = C.CDLL(str(SCENES_LIB_PATH))
        lib.rtbs_build.restype = C.c_void_p
        lib.rtbs_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        lib.rtbs_desc.restype = C.POINTER(RtbSceneDesc)
        lib.rtbs_desc.argtypes = [C.c_void_p]
        lib.rtbs_free.argtypes = [C.c_void_p]
        lib.rtbs_free.restype = None
        lib.rtbs_last_error.restype = C.c_char_p
        _scenes_lib = lib
    return _scenes_lib
