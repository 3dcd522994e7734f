"""In-tree build of the native pieces (called by __graft_entry__.build()).

  librtb200.so       CUDA kernels (sm_100a) + C ABI               <- the product
  librtb_scenes.so   C++ host mirror of the reference scene API + BASELINE configs (test/bench input)

nvcc cross-compiles without a GPU.  The .so files are git-ignored but travel to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(map(str, cmd)), file=sys.stderr)
    r = subprocess.run(list(map(str, cmd)), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(map(str, cmd))}\n{r.stdout}\n{r.stderr}")
    return r


def build_cuda_library(force: bool = False, verbose: bool = False, extra_flags=()) -> Path:
    out = PKG / "librtb200.so"
    srcs = [CSRC / "kernels.cu", CSRC / "wavefront.cu", CSRC / "api.cpp", CSRC / "flatten.cpp"]
    deps = srcs + [CSRC / "rtb_device.cuh", CSRC / "device_scene.h", CSRC / "kernels.h", CSRC / "flatten.h",
                   PKG.parent / "include" / "rtb200.h"]
    if force or _stale(out, deps):
        _run([_nvcc(), *NVCC_FLAGS, *extra_flags, "-shared", "-o", out, *srcs], verbose)
    return out


def build_scenes_library(force: bool = False, verbose: bool = False) -> Path:
    out = PKG / "librtb_scenes.so"
    src = PKG / "host" / "scenes.cpp"
    deps = [src, PKG / "host" / "rtb" / "scene.hpp", PKG.parent / "include" / "rtb200.h"]
    if force or _stale(out, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", out, src], verbose)
    return out


def build_example(force: bool = False, verbose: bool = False) -> Path:
    """host/example_cornell: a main.rs-style scene on the C++ mirror, linked against librtb200.so."""
    out = PKG / "example_cornell"
    src = PKG / "host" / "example_cornell.cpp"
    deps = [src, PKG / "host" / "rtb" / "scene.hpp", PKG / "host" / "rtb" / "render.hpp", PKG / "librtb200.so"]
    if force or _stale(out, deps):
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-o", out, src, f"-L{PKG}", "-lrtb200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN"], verbose)
    return out


def build_all(force: bool = False, verbose: bool = False):
    return [build_cuda_library(force, verbose), build_scenes_library(force, verbose), build_example(force, verbose)]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose=True):
        print(p)
