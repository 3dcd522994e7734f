"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/rtb200.h declares,
and its host-side logic (validation, flatten, error reporting) behaves.  No compute calls without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from surely_raytracing_b200 import capi
from surely_raytracing_b200.scenes import BuiltScene
from tests.conftest import has_gpu

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    header = (ROOT / "include" / "rtb200.h").read_text()
    declared = sorted(set(re.findall(r"\b(rtb_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(capi.EXPORTS), "capi.EXPORTS is out of sync with include/rtb200.h"
    for name in declared:
        assert hasattr(lib, name), f"librtb200.so does not export {name}"
    assert lib.rtb_version() == capi.RTB_ABI_VERSION


def test_struct_layouts_match_the_header():
    # sizes computed by hand from include/rtb200.h (LP64)
    assert C.sizeof(capi.RtbObject) == 16 + 80
    assert C.sizeof(capi.RtbMaterial) == 8 + 24 + 8
    assert C.sizeof(capi.RtbTexture) == 16 + 24 + 8
    assert C.sizeof(capi.RtbImage) == 16
    assert C.sizeof(capi.RtbPerlin) == 256 * 24 + 3 * 256 * 4
    assert C.sizeof(capi.RtbCamera) == 8 + 16 + 8 + 72 + 16 + 24
    assert C.sizeof(capi.RtbRay) == 64 and C.sizeof(capi.RtbHit) == 16 + 8 + 48 + 16
    assert C.sizeof(capi.RtbRenderParams) == 24 and C.sizeof(capi.RtbStats) == 64 + 16 + 24
    assert C.sizeof(capi.RtbSun) == 56
    # RtbSceneDesc: 16 header + 5 x (pointer, count|index pairs) ... checked against the compiler below
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "rtb200.h"
        int main(void) { printf("%zu %zu %zu %zu\\n", sizeof(RtbSceneDesc), sizeof(RtbStats), sizeof(RtbSceneInfo), sizeof(RtbCamera)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as td:
        (Path(td) / "s.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", f"{td}/s", f"{td}/s.c"], check=True)   # the header is plain C
        sizes = list(map(int, subprocess.run([f"{td}/s"], capture_output=True, text=True, check=True).stdout.split()))
    assert sizes == [C.sizeof(capi.RtbSceneDesc), C.sizeof(capi.RtbStats), C.sizeof(capi.RtbSceneInfo), C.sizeof(capi.RtbCamera)]


def _create(desc):
    lib = capi.load_library()
    h = C.c_void_p()
    rc = lib.rtb_scene_create(desc, 0, C.byref(h))
    msg = lib.rtb_last_error().decode()
    if rc == 0:
        lib.rtb_scene_destroy(h)
    return rc, msg


def test_invalid_descriptions_are_rejected_with_a_message():
    """The reference panics on malformed input (expect/panic!); the ABI returns a code + message instead.
    Validation runs on the host before any CUDA call, so this works without a GPU."""
    b = BuiltScene("c2", width=16, spp=4)
    d = b.desc.contents
    saved = d.world
    d.world = d.n_objects + 5
    rc, msg = _create(b.desc)
    assert rc == -1 and "world" in msg
    d.world = saved

    saved_v = d.abi_version
    d.abi_version = 99
    rc, msg = _create(b.desc)
    assert rc == -1 and "abi" in msg
    d.abi_version = saved_v

    mat = d.objects[1].material
    assert d.objects[1].kind == 1  # first wall quad
    d.objects[1].material = 1000
    rc, msg = _create(b.desc)
    assert rc == -1 and "material" in msg
    d.objects[1].material = mat

    # a shared object (DAG) is refused: flatten must emit one node per occurrence
    ch0, ch1 = d.children[0], d.children[1]
    d.children[1] = ch0
    rc, msg = _create(b.desc)
    assert rc == -1 and "twice" in msg
    d.children[1] = ch1

    w = d.camera.image_width
    d.camera.image_width = 0
    rc, msg = _create(b.desc)
    assert rc == -1 and "camera" in msg
    d.camera.image_width = w


def test_no_device_means_an_error_not_a_fallback():
    if has_gpu():
        pytest.skip("a GPU is visible")
    b = BuiltScene("c2", width=16, spp=4)  # keep the owner of the description alive
    rc, msg = _create(b.desc)
    assert rc == -4 and "no CPU fallback" in msg


def test_every_config_flattens_on_the_host():
    """flatten + BVH build of all BASELINE configs through the host compiler used by librtb200.so
    (exercised here through the emulation harness, which links the same flatten.cpp)."""
    from tests.emu.emu_lib import EmuScene
    expect = {"c1": None, "c2": (18, 0, 0, 1), "c3": (6, 12, 2, 0), "c4": (3407, 2, 2, 0), "c5": (13, 0, 0, 2)}
    for cfg, exp in expect.items():
        e = EmuScene(BuiltScene(cfg, width=32, spp=4))
        i = e.info
        assert i.bvh_depth < 40 and i.n_bvh_nodes >= 1
        if exp:
            assert (i.n_surface_prims, i.n_boundary_prims, i.n_media, i.n_lights) == exp, cfg
        else:
            assert 400 < i.n_surface_prims <= 488


def test_no_kernel_carries_a_uniform_register_across_an_escaping_loop_exit():
    """tools/sass_lint.py over the shipped library.  Round 1's "illegal memory access" was ptxas keeping a loop-invariant
    address in a (per-warp) uniform register across the divergent traversal loop and reusing that register in code the
    early leavers of the loop could reach (tests/golden/sass_fault_r01.txt is that kernel's listing: the lint must flag
    it); every divergent loop now ends in an explicit reconvergence, and no shipped kernel may show the pattern."""
    import subprocess
    import sys
    lint = str(ROOT / "tools" / "sass_lint.py")
    bad = subprocess.run([sys.executable, lint, str(ROOT / "tests" / "golden" / "sass_fault_r01.txt")], capture_output=True, text=True)
    assert bad.returncode == 1 and "UR5" in bad.stdout, bad.stdout
    good = subprocess.run([sys.executable, lint, str(capi.LIB_PATH)], capture_output=True, text=True)
    assert good.returncode == 0 and good.stdout.count("ok ") >= 40, good.stdout[-2000:]
