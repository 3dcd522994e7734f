#!/usr/bin/env python
"""Build A/B variants of librtb200.so (compile-time -D flags) into surely_raytracing_b200/variants/
for tuning runs on the GPU box (selected with RTB200_LIB).  Development tool; not part of the product."""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from surely_raytracing_b200 import build as B  # noqa: E402

OUT = B.PKG / "variants"


def build(name, flags):
    OUT.mkdir(exist_ok=True)
    out = OUT / f"librtb200_{name}.so"
    srcs = [B.CSRC / "kernels.cu", B.CSRC / "wavefront.cu", B.CSRC / "api.cpp", B.CSRC / "flatten.cpp"]
    cmd = [B._nvcc(), *B.NVCC_FLAGS, *flags, "-shared", "-o", str(out), *map(str, srcs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(name, "FAILED\n", r.stderr[-2000:])
    return name, r.returncode


if __name__ == "__main__":
    # usage: variants.py name1="-DA=1 -DB=2" name2="..."
    jobs = []
    for a in sys.argv[1:]:
        n, _, f = a.partition("=")
        jobs.append((n, f.split()))
    with ThreadPoolExecutor(max_workers=6) as ex:
        for n, rc in ex.map(lambda j: build(*j), jobs):
            print(n, "ok" if rc == 0 else "FAILED")
