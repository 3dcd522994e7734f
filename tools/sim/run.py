#!/usr/bin/env python
"""Drive tools/sim/sim.cpp: collect realistic ray queues of c4 on the CPU and rank extend policies.
Development aid only (no GPU, no product code path)."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from surely_raytracing_b200 import BuiltScene  # noqa: E402
from surely_raytracing_b200 import capi  # noqa: E402

NO_BOX = capi.RTB_FLAG_NO_BOX_LEAVES  # the prototypes scan primitives one by one through their DPre records


HERE = Path(__file__).resolve().parent
LIB = HERE / "libsim.so"


def build():
    srcs = [HERE / "sim.cpp", ROOT / "surely_raytracing_b200/csrc/flatten.cpp"]
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                    "-o", str(LIB), *map(str, srcs)], check=True)


class Policy(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("threshold", "leaf_slots", "stale_cull", "break_mode", "break_count", "n_warps",
                                       "cost_inner", "cost_quad", "cost_sphere", "cost_fetch", "cost_outer")]


class SimOut(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("rays", "inner_trips", "inner_lane_sum", "leaf_rounds", "leaf_lane_sum", "node_visits",
                                          "prim_tests", "fetches", "outer_trips", "cost", "stale_skipped", "distinct_nodes")]


QRAY = np.dtype([("o", "<f8", 3), ("d", "<f4", 3), ("time", "<f4")])


def main():
    build()
    lib = C.CDLL(str(LIB))
    lib.emu_scene_create.restype = C.c_void_p
    lib.sim_collect.restype = C.c_longlong
    width = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    b = BuiltScene("c4", width=width, flags=NO_BOX)
    h = C.c_void_p(lib.emu_scene_create(b.desc))
    cache = HERE / f"rays_w{width}_c{cap}.npy"
    if cache.exists():
        rays = np.load(cache)
    else:
        it_lo, it_hi = 8, 12
        buf = np.zeros(cap * (it_hi - it_lo), dtype=QRAY)
        offs = np.zeros(it_hi - it_lo + 1, dtype=np.int64)
        n = lib.sim_collect(h, cap, C.c_longlong(0), C.c_longlong(64), it_lo, it_hi, buf.ctypes.data_as(C.c_void_p),
                            C.c_longlong(len(buf)), offs.ctypes.data_as(C.c_void_p))
        rays = buf[:n]
        np.save(cache, rays)
    print("rays", len(rays))
    import os
    base = dict(threshold=28, leaf_slots=1, stale_cull=0, break_mode=0, break_count=0, n_warps=64, cost_inner=80,
                cost_quad=int(os.environ.get("CQ", 90)), cost_sphere=int(os.environ.get("CS", 110)), cost_fetch=70,
                cost_outer=int(os.environ.get("CO", 25)))
    variants = [("shipped", {})]
    for bc in (10, 12, 14, 16, 18, 20, 24):
        variants.append((f"left>={bc}", dict(break_mode=3, break_count=bc)))
    for bc in (12, 16, 20):
        variants.append((f"left>={bc} stale", dict(break_mode=3, break_count=bc, stale_cull=1)))
        variants.append((f"left>={bc} slots2", dict(break_mode=3, break_count=bc, leaf_slots=2)))
        variants.append((f"left>={bc} thr24", dict(break_mode=3, break_count=bc, threshold=24)))
        variants.append((f"left>={bc} thr32", dict(break_mode=3, break_count=bc, threshold=32)))
        variants.append((f"left>={bc} thr20 stale", dict(break_mode=3, break_count=bc, threshold=20, stale_cull=1)))
    print(f"{'policy':22s} {'cost/ray':>9s} {'inner/ray':>9s} {'lanes_in':>8s} {'leafrnd/ray':>11s} {'lanes_lf':>8s} "
          f"{'nodes/ray':>9s} {'prims/ray':>9s} {'stale/ray':>9s}")
    for name, kv in variants:
        pol = Policy(**{**base, **kv})
        out = SimOut()
        lib.sim_extend(h, rays.ctypes.data_as(C.c_void_p), C.c_longlong(len(rays)), C.byref(pol), C.byref(out))
        r = out.rays
        print(f"{name:22s} {out.cost / r:9.2f} {out.inner_trips / r:9.3f} {out.inner_lane_sum / max(out.inner_trips, 1):8.2f} "
              f"{out.leaf_rounds / r:11.3f} {out.leaf_lane_sum / max(out.leaf_rounds, 1):8.2f} {out.node_visits / r:9.2f} "
              f"{out.prim_tests / r:9.3f} {out.stale_skipped / r:9.3f}  distinct/trip {out.distinct_nodes / out.inner_trips:6.2f}")




def sort_experiment():
    """How much would a coherence sort of the queue buy?  (upper bounds: the sort itself is not costed)"""
    lib = C.CDLL(str(LIB))
    lib.emu_scene_create.restype = C.c_void_p
    width, cap = 160, 32768
    b = BuiltScene("c4", width=width, flags=NO_BOX)
    h = C.c_void_p(lib.emu_scene_create(b.desc))
    rays = np.load(HERE / f"rays_w{width}_c{cap}.npy")
    base = dict(threshold=28, leaf_slots=1, stale_cull=0, break_mode=0, break_count=0, n_warps=64, cost_inner=80,
                cost_quad=90, cost_sphere=110, cost_fetch=70, cost_outer=25)

    def run(name, r):
        r = np.ascontiguousarray(r)
        pol = Policy(**base)
        out = SimOut()
        lib.sim_extend(h, r.ctypes.data_as(C.c_void_p), C.c_longlong(len(r)), C.byref(pol), C.byref(out))
        print(f"{name:28s} cost/ray {out.cost / out.rays:7.2f} inner/ray {out.inner_trips / out.rays:6.3f} lanes {out.inner_lane_sum / out.inner_trips:6.2f} "
              f"leaf rounds/ray {out.leaf_rounds / out.rays:6.3f}")

    def morton(o, bits):
        lo, hi = o.min(axis=0), o.max(axis=0)
        q = np.clip(((o - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
        code = np.zeros(len(o), dtype=np.int64)
        for bit in range(bits):
            for a in range(3):
                code |= ((q[:, a] >> bit) & 1) << (3 * bit + a)
        return code

    run("queue order", rays)
    chunks = [rays[i:i + cap] for i in range(0, len(rays), cap)]
    for bits in (2, 3, 5, 8):
        out = []
        for ch in chunks:
            d = ch["d"]
            octant = (d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2)
            key = (morton(np.clip(ch["o"], -1200, 1200), bits) << 3) | octant
            out.append(ch[np.argsort(key, kind="stable")])
        run(f"sort morton{bits}+octant", np.concatenate(out))
    out = []
    for ch in chunks:
        d = ch["d"]
        octant = (d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2)
        out.append(ch[np.argsort(octant, kind="stable")])
    run("sort octant only", np.concatenate(out))
    rng = np.random.default_rng(1)
    run("shuffled", np.concatenate([ch[rng.permutation(len(ch))] for ch in chunks]))


if len(sys.argv) > 1 and sys.argv[1] == "sort":
    sort_experiment()
elif len(sys.argv) > 1 and sys.argv[1] in ("checkq", "bvh4", "prefilter", "candidates"):
    pass
elif __name__ == "__main__":
    main()


def check_q(width=160, cap=32768):
    build()
    lib = C.CDLL(str(LIB))
    lib.emu_scene_create.restype = C.c_void_p
    for cfg in ("c4", "c1", "c2", "c3", "c5"):
        b = BuiltScene(cfg, width=width, flags=NO_BOX)
        h = C.c_void_p(lib.emu_scene_create(b.desc))
        if cfg == "c4":
            rays = np.load(HERE / f"rays_w{width}_c{cap}.npy")
        else:
            buf = np.zeros(cap * 3, dtype=QRAY)
            offs = np.zeros(4, dtype=np.int64)
            n = lib.sim_collect(h, cap, C.c_longlong(0), C.c_longlong(16), 2, 5, buf.ctypes.data_as(C.c_void_p),
                                C.c_longlong(len(buf)), offs.ctypes.data_as(C.c_void_p))
            rays = buf[:n]
        # plus adversarial rays: axis-parallel directions, far origins
        rng = np.random.default_rng(7)
        extra = rays[rng.integers(0, len(rays), 20000)].copy()
        extra["d"][:5000, 0] = 0.0
        extra["d"][5000:10000, 1] = 0.0
        extra["o"][10000:15000] *= 40.0
        extra["o"][15000:] += rng.normal(0, 3000, (5000, 3))
        allr = np.ascontiguousarray(np.concatenate([rays, extra]))
        out = np.zeros(6)
        lib.emu_check_qnodes(h, allr.ctypes.data_as(C.c_void_p), C.c_longlong(len(allr)), C.c_longlong(7), out.ctypes.data_as(C.c_void_p))
        print(f"{cfg}: rays {len(allr)}  node visits/ray q {out[0]:.3f} fp32 {out[1]:.3f}  mismatches vs fp32-tree {int(out[2])}  "
              f"vs brute force {int(out[3])} of {int(out[4])}  unculled rays {int(out[5])}")
        lib.emu_check_nodes4(h, allr.ctypes.data_as(C.c_void_p), C.c_longlong(len(allr)), C.c_longlong(7), out.ctypes.data_as(C.c_void_p))
        print(f"{cfg}: bvh4 visits/ray {out[0]:.3f} (bvh2 {out[1]:.3f})  mismatches vs bvh2 {int(out[2])}  vs brute force {int(out[3])} of {int(out[4])}  "
              f"max stack {int(out[5])}")


if len(sys.argv) > 1 and sys.argv[1] == "checkq":
    check_q()


def bvh4():
    build()
    lib = C.CDLL(str(LIB))
    lib.emu_scene_create.restype = C.c_void_p
    b = BuiltScene("c4", width=160, flags=NO_BOX)
    h = C.c_void_p(lib.emu_scene_create(b.desc))
    rays = np.load(HERE / "rays_w160_c32768.npy")
    base = dict(threshold=28, leaf_slots=1, stale_cull=0, break_mode=0, break_count=0, n_warps=64, cost_inner=80,
                cost_quad=90, cost_sphere=110, cost_fetch=70, cost_outer=25)
    for name, fn, kv in (("bvh2 shipped", lib.sim_extend, {}), ("bvh2 left>=16", lib.sim_extend, dict(break_mode=3, break_count=16)),
                         ("bvh4", lib.sim_extend4, {}), ("bvh4 left>=16", lib.sim_extend4, dict(break_mode=3, break_count=16)),
                         ("bvh4 left>=12", lib.sim_extend4, dict(break_mode=3, break_count=12)),
                         ("bvh4 l16 nearest-only", lib.sim_extend4, dict(break_mode=3, break_count=16, stale_cull=1))):
        pol = Policy(**{**base, **kv})
        out = SimOut()
        fn(h, rays.ctypes.data_as(C.c_void_p), C.c_longlong(len(rays)), C.byref(pol), C.byref(out))
        r = out.rays
        print(f"{name:16s} trips/ray {out.inner_trips / r:6.3f} lanes {out.inner_lane_sum / out.inner_trips:6.2f} visits/ray {out.node_visits / r:6.2f} "
              f"leaf rounds/ray {out.leaf_rounds / r:6.3f} leaf lanes {out.leaf_lane_sum / out.leaf_rounds:5.2f} prims/ray {out.prim_tests / r:5.3f} "
              f"distinct/trip {out.distinct_nodes / out.inner_trips:5.2f}")


if len(sys.argv) > 1 and sys.argv[1] == "bvh4":
    bvh4()


def prefilter(width=160, cap=32768):
    build()
    lib = C.CDLL(str(LIB))
    lib.emu_scene_create.restype = C.c_void_p
    lib.sim_collect.restype = C.c_longlong
    for cfg in ("c4", "c1", "c2", "c3", "c5"):
        b = BuiltScene(cfg, width=width, flags=NO_BOX)
        h = C.c_void_p(lib.emu_scene_create(b.desc))
        if cfg == "c4" and (HERE / f"rays_w{width}_c{cap}.npy").exists():
            rays = np.load(HERE / f"rays_w{width}_c{cap}.npy")
        else:
            buf = np.zeros(cap * 3, dtype=QRAY)
            offs = np.zeros(4, dtype=np.int64)
            n = lib.sim_collect(h, cap, C.c_longlong(0), C.c_longlong(16), 2, 5, buf.ctypes.data_as(C.c_void_p),
                                C.c_longlong(len(buf)), offs.ctypes.data_as(C.c_void_p))
            rays = buf[:n]
        out = np.zeros(8)
        r = np.ascontiguousarray(rays)
        lib.sim_prefilter(h, r.ctypes.data_as(C.c_void_p), C.c_longlong(len(r)), out.ctypes.data_as(C.c_void_p))
        t = out[0]
        print(f"{cfg}: {int(t)} leaf tests of {len(r)} rays  certain miss {100 * out[1] / t:5.1f} %  certain hit {100 * out[2] / t:5.1f} %  "
              f"uncertain {100 * out[3] / t:5.2f} %  exact hits {100 * out[5] / t:5.1f} %  VIOLATIONS {int(out[4])}  "
              f"mean relative width of the t bounds {out[6] / max(out[2], 1):.2e}")


if len(sys.argv) > 1 and sys.argv[1] == "prefilter":
    prefilter()


def candidates(width=160, cap=32768):
    build()
    lib = C.CDLL(str(LIB))
    lib.emu_scene_create.restype = C.c_void_p
    lib.sim_collect.restype = C.c_longlong
    for cfg in ("c4", "c1", "c2", "c3", "c5"):
        b = BuiltScene(cfg, width=width, flags=NO_BOX)
        h = C.c_void_p(lib.emu_scene_create(b.desc))
        if cfg == "c4" and (HERE / f"rays_w{width}_c{cap}.npy").exists():
            rays = np.load(HERE / f"rays_w{width}_c{cap}.npy")
        else:
            buf = np.zeros(cap * 3, dtype=QRAY)
            offs = np.zeros(4, dtype=np.int64)
            n = lib.sim_collect(h, cap, C.c_longlong(0), C.c_longlong(16), 2, 5, buf.ctypes.data_as(C.c_void_p),
                                C.c_longlong(len(buf)), offs.ctypes.data_as(C.c_void_p))
            rays = buf[:n]
        r = np.ascontiguousarray(rays)
        for K in (2,):
            out = np.zeros(8)
            lib.sim_candidates(h, r.ctypes.data_as(C.c_void_p), C.c_longlong(len(r)), K, out.ctypes.data_as(C.c_void_p))
            nr = out[0]
            print(f"{cfg} K={K}: rays {int(nr)}  MISMATCHES {int(out[1])}  fallbacks {100 * out[2] / nr:6.3f} %  candidates/ray {out[3] / max(nr - out[2], 1):5.3f}  "
                  f"rays with 2 candidates {100 * out[6] / nr:5.2f} %  node visits/ray {out[4] / nr:6.2f} (exact scheme {out[5] / nr:6.2f})  prefilter tests/ray {out[7] / nr:5.3f}")


if len(sys.argv) > 1 and sys.argv[1] == "candidates":
    candidates()
