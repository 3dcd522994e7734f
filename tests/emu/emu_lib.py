"""ctypes loader for tests/emu/libemu.so (device header compiled for the host; development aid)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from surely_raytracing_b200 import capi

EMU_DIR = Path(__file__).resolve().parent
LIB = EMU_DIR / "libemu.so"
ROOT = EMU_DIR.parent.parent


def build(force=False):
    srcs = [EMU_DIR / "emu.cpp", ROOT / "surely_raytracing_b200/csrc/flatten.cpp", ROOT / "surely_raytracing_b200/csrc/rtb_device.cuh",
            ROOT / "surely_raytracing_b200/csrc/device_scene.h", ROOT / "surely_raytracing_b200/csrc/flatten.h"]
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", str(LIB),
                        str(srcs[0]), str(srcs[1])], check=True, capture_output=True)


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        vp, i64 = C.c_void_p, C.c_longlong
        lib.emu_last_error.restype = C.c_char_p
        lib.emu_scene_create.restype = vp
        lib.emu_scene_create.argtypes = [C.POINTER(capi.RtbSceneDesc)]
        lib.emu_scene_destroy.argtypes = [vp]
        lib.emu_scene_info.argtypes = [vp, C.POINTER(capi.RtbSceneInfo)]
        lib.emu_render.argtypes = [vp, i64, i64, vp, vp]
        lib.emu_trace.argtypes = [vp, vp, i64, C.c_uint, vp]
        lib.emu_trace_candidates.argtypes = [vp, vp, i64, vp, vp]
        lib.emu_medium_interval.argtypes = [vp, C.c_int, vp, i64, vp, vp]
        lib.emu_eval_texture.argtypes = [vp, C.c_int, vp, i64, vp]
        lib.emu_eval_light_pdf.argtypes = [vp, vp, i64, vp]
        lib.emu_spec_bits.argtypes = [vp]
        lib.emu_medium_flags.argtypes = [vp, C.c_int]
        lib.emu_defer_ok.argtypes = [vp]
        lib.emu_check_leaf_refs.argtypes = [vp]
        lib.emu_check_qnodes.argtypes = [vp, vp, i64, i64, vp]
        lib.emu_check_nodes4.argtypes = [vp, vp, i64, i64, vp]
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class EmuScene:
    def __init__(self, built):
        self._lib = load()
        self._h = self._lib.emu_scene_create(built.desc)
        if not self._h:
            raise RuntimeError(self._lib.emu_last_error().decode())
        self._built = built
        self.info = capi.RtbSceneInfo()
        self._lib.emu_scene_info(self._h, C.byref(self.info))

    def render(self, sample_begin=0, sample_end=None):
        if sample_end is None:
            sample_end = self.info.spp_used
        out = np.zeros((self.info.image_height, self.info.image_width, 3))
        st = np.zeros(6, dtype=np.uint64)
        self._lib.emu_render(self._h, sample_begin, sample_end, _ptr(out), _ptr(st))
        names = ["paths", "segments", "node_visits", "prim_tests", "medium_probes", "nonfinite_samples"]
        return out, dict(zip(names, map(int, st)))

    def trace(self, rays, flags=0):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=capi.HIT_DTYPE)
        self._lib.emu_trace(self._h, _ptr(rays), len(rays), flags, _ptr(hits))
        return hits

    def trace_candidates(self, rays):
        """closest hits through the candidate scheme (conservative fp32 classification, exact resolution of the
        survivors); returns (hits, {overflows, resolved, two_candidates})"""
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=capi.HIT_DTYPE)
        c = np.zeros(3, dtype=np.uint64)
        if self._lib.emu_trace_candidates(self._h, _ptr(rays), len(rays), _ptr(hits), _ptr(c)) != 0:
            raise RuntimeError(self._lib.emu_last_error().decode())
        return hits, {"overflows": int(c[0]), "resolved": int(c[1]), "two_candidates": int(c[2])}

    def medium_interval(self, medium, rays):
        rays = np.ascontiguousarray(rays, dtype=capi.RAY_DTYPE)
        t0 = np.zeros(len(rays)); t1 = np.zeros(len(rays))
        self._lib.emu_medium_interval(self._h, medium, _ptr(rays), len(rays), _ptr(t0), _ptr(t1))
        return t0, t1

    def eval_texture(self, texture, uvp):
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((len(uvp), 3))
        self._lib.emu_eval_texture(self._h, texture, _ptr(uvp), len(uvp), _ptr(out))
        return out

    def spec_bits(self):
        return int(self._lib.emu_spec_bits(self._h))

    def defer_ok(self):
        return int(self._lib.emu_defer_ok(self._h))

    def leaf_ref_violations(self):
        return int(self._lib.emu_check_leaf_refs(self._h))

    QRAY_DTYPE = np.dtype([("o", "<f8", 3), ("d", "<f4", 3), ("time", "<f4")])  # a ray as the wavefront queues store it

    def check_trees(self, rays, brute_every=5):
        """Closest hits through the 32-byte quantised nodes and through the collapsed BVH4 against the fp32
        BVH2 and a brute-force scan.  Returns two dicts of counts."""
        q = np.zeros(len(rays), dtype=self.QRAY_DTYPE)
        q["o"] = rays["origin"]; q["d"] = rays["direction"]; q["time"] = rays["time"]
        out = {}
        for name, fn in (("qnodes", self._lib.emu_check_qnodes), ("nodes4", self._lib.emu_check_nodes4)):
            r = np.zeros(6)
            fn(self._h, _ptr(q), len(q), brute_every, _ptr(r))
            out[name] = {"visits": r[0], "visits_bvh2": r[1], "mismatch_bvh2": int(r[2]), "mismatch_brute": int(r[3]),
                         "n_brute": int(r[4]), "extra": int(r[5])}
        return out

    def eval_light_pdf(self, od):
        od = np.ascontiguousarray(od, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(od))
        self._lib.emu_eval_light_pdf(self._h, _ptr(od), len(od), _ptr(out))
        return out

    def __del__(self):
        try:
            if self._h:
                self._lib.emu_scene_destroy(self._h)
        except Exception:
            pass
