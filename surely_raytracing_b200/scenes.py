"""Flat scene descriptions of BASELINE.json's configs (built by host/scenes.cpp on the C++ mirror of
the reference's construction API).  `BuiltScene(name)` owns one description; `.desc` is what
rtb_scene_create (and the test oracle) consume."""
from __future__ import annotations

import ctypes as C

from . import capi

DEFAULT_SEED = 20240001
CONFIGS = {
    "c1": "Book1 final random-spheres scene 400x225 @ 50 spp (main.rs:135-210)",
    "c2": "Cornell box, two boxes, lights=[quad] 600x600 @ 1000 spp (main.rs:417-512 with box2)",
    "c3": "Cornell smoke 600x600 @ 2000 spp depth 10 (main.rs:514-601)",
    "c4": "Book2 final scene 800x800 @ 10000 spp depth 40 (main.rs:603-712)",
    "c5": "Book3 mixed-PDF Cornell box at HEAD, lights=[quad, sphere] 600x600 @ 1000 spp (main.rs:417-512)",
}


class BuiltScene:
    def __init__(self, name: str, width: int = 0, spp: int = 0, depth: int = 0, seed: int = DEFAULT_SEED,
                 variant: int = 0, flags: int = 0):
        self._lib = capi.load_scenes_library()
        self.name = name
        self._h = self._lib.rtbs_build(name.encode(), width, spp, depth, seed, variant, flags)
        if not self._h:
            raise capi.RtbError(self._lib.rtbs_last_error().decode())
        self.desc = self._lib.rtbs_desc(self._h)

    @property
    def camera(self) -> capi.RtbCamera:
        return self.desc.contents.camera

    def close(self):
        if self._h:
            self._lib.rtbs_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
