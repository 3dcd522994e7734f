#!/usr/bin/env python
"""bench.py -- Mpaths/s (camera samples/s) of the hot path on BASELINE.json's headline config:
the book-2 final scene (c4) at 800x800, depth 40, on N B200s of one box.

A "step" is one progressive pass: one rtb_render call over 10 rows of the 100x100 stratum grid
(1000 strata for every one of the 640,000 pixels = 640 M paths; --rows-per-step).
  --scaling weak (default): every GPU renders its own 640 M-path pass per step (row blocks dealt round-robin);
  --scaling strong: the 640 M paths of a step are split N ways (contiguous stratum slices), so
      `--scaling strong --steps 10` is THE target job: final_scene(800, 10000, 40) = 6.4 G paths on N GPUs.
Each rank accumulates into its own int64 fixed-point buffer and ONE NCCL sum-reduce at the end of the timed region
delivers the image to rank 0; when the passes cover whole stratum grids the reduced image is accepted
against the committed oracle render (tests/golden) in the epilogue.

Printed: ONE JSON line (see the task contract) with `value` (device-resident), `e2e` (through the C-ABI calls with
host buffers: scene upload + render + reduce + read-back every step), `roofline`, `cpu_baseline`, `clocks`,
`gpu_launches`, `configs` (the other BASELINE configs at full size, N = 1).  `--impl reference` times the CPU
restatement of the reference (the oracle: the Rust reference cannot be built in this image) on the host cores.
`--workload c5 --curve` writes the variance-vs-time curve of BASELINE config 5.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = "c4"
WORKLOAD_DESC = ("book-2 final scene (final_scene, reference src/main.rs:603-712) 800x800 depth 40, lights = empty "
                 "(what main.rs passes); step = 10 rows (1000 strata) of the 100x100 stratum grid")
L2_FLUSH_BYTES = 256 << 20
ROWS_PER_STEP = 10

# ALGORITHMIC lane-instruction constants per call: SURVEY.md section 8(a)'s own table (a3, a6, a8, a9, a11, a13, a22),
# written down before any kernel existed and FROZEN -- they describe the reference's work, not this implementation's
# SASS, so `roofline.frac` moves only with speed and with the traversal's own call counts:
#   box      19   Aabb::hit per box with inv_d precomputed (a6); a BVH2 node visit tests two boxes
#   prim     31   one leaf primitive test: 0.57 x (quad full 47 / sphere hit 40) + 0.43 x (quad reject 12 / sphere miss 17) (a8, a9)
#   medium   97   ConstantMedium::hit = two boundary sphere hits + 17 (a11), per medium and segment
#   shade   145   one Lambertian bounce incl. ONB, mixture sample, pdfs (a13-a16); + 63 per listed light (a16)
#   philox   60   Philox4x32-10 (10 rounds x 6), two calls per segment (a22), one per primary ray
#   get_ray  25   stratified jittered primary ray (a3)
I_CONST = {"box": 19.0, "prim": 31.0, "medium": 97.0, "shade": 145.0, "light": 63.0, "philox": 60.0, "get_ray": 25.0}
PEAK_LANE_INSTR_PER_CLK_PER_SM = 128
# ncu counters of the FINAL binary (profiles/r02_*): executed thread-instructions per segment of the two hot kernels,
# their lanes per warp instruction, and DRAM bytes per segment over every k_wf_* launch of the bench command.
NCU = {"source": "profiles/r02_launch_summary.txt (ncu launch list of this command on the final binary; tools/ncu_constants.py)",
       "thread_instr_per_segment": {"k_wf_extend": 1069.6, "k_wf_shade": 860.1},
       "lanes": {"k_wf_extend": 16.28, "k_wf_shade": 21.54},
       "dram_bytes_per_segment_c4": 199.0}


def algorithmic_instr(st, n_lights):
    """(I_path, I_extend per segment, I_shade per segment) from the traversal's own counters of a counted pass"""
    seg = st["segments"]
    i_extend = (st["node_visits"] * 2 * I_CONST["box"] + st["prim_tests"] * I_CONST["prim"]) / seg
    i_shade = (st["medium_probes"] * I_CONST["medium"]) / seg + I_CONST["shade"] + n_lights * I_CONST["light"] + 2 * I_CONST["philox"]
    i_path = I_CONST["get_ray"] + I_CONST["philox"] + (i_extend + i_shade) * seg / st["paths"]
    return i_path, i_extend, i_shade


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pipeline", default="default", choices=["default", "mega", "wavefront"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--width", type=int, default=0, help="override the image width (parity/debug only)")
    ap.add_argument("--rows-per-step", type=int, default=ROWS_PER_STEP,
                    help="rows of the sqrt x sqrt stratum grid rendered per step (one rtb_render call)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1/c2/c3/c5 block")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--option", action="append", default=[], help="RTB_OPT id=value (tuning / A-B runs), e.g. 2=1 for exact leaves")
    ap.add_argument("--flags", type=lambda v: int(v, 0), default=0, help="RTB_FLAG_* bits of the scene description (builder arms)")
    ap.add_argument("--curve", action="store_true", help="variance-vs-time curve (BASELINE config 5): RMSE vs the oracle at growing spp")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference arm: the reference's own CPU implementation of the path.  The Rust crate cannot be
    compiled here (no cargo/rustc, crates not vendored), so this is the oracle port (reference sampler
    mode: sequential stream, rejection loops, recursive ray_color, reference-shaped BVH) with all host
    threads, on the same config / metric.  Each step is a bounded sample: `spp_per_step` strata of c4."""
    if rank != 0:
        return
    from oracle import orc
    from surely_raytracing_b200.scenes import BuiltScene
    b = BuiltScene(args.workload, width=args.width)
    o = orc.OracleScene(b, use_bvh=True)
    n_px = o.info.image_width * o.info.image_height
    threads = os.cpu_count() or 1
    # calibrate: one stratum
    t0 = time.perf_counter()
    o.render(0, 1, sampler=orc.SAMPLER_REF, threads=threads)
    dt = time.perf_counter() - t0
    budget = 120.0 / max(1, args.steps + args.warmup)          # whole run inside ~2 minutes
    spp_step = max(1, min(o.info.sqrt_spp, int(budget / max(dt, 1e-3))))
    for k in range(args.warmup):
        o.render(k * spp_step, (k + 1) * spp_step, sampler=orc.SAMPLER_REF, threads=threads)
    t0 = time.perf_counter()
    segs = 0
    for k in range(args.steps):
        lo = ((args.warmup + k) * spp_step) % (o.info.spp_used - spp_step + 1)
        _, st = o.render(lo, lo + spp_step, sampler=orc.SAMPLER_REF, threads=threads)
        segs += st["segments"]
    dt = time.perf_counter() - t0
    paths = n_px * spp_step * args.steps
    value = paths / dt / 1e6
    sample = f"{spp_step} strata x {n_px} pixels per step ({paths} paths in {dt:.1f} s)"
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC if args.workload == WORKLOAD else args.workload, "sample": sample,
                   "note": "CPU restatement (oracle port) of the Rust reference: cargo/rustc absent in this image"},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "segments_per_path": segs / max(1, paths), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def cpu_baseline(workload, width, target_seconds):
    from oracle import orc
    from surely_raytracing_b200.scenes import BuiltScene
    b = BuiltScene(workload, width=width)
    o = orc.OracleScene(b, use_bvh=True)
    n_px = o.info.image_width * o.info.image_height
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    o.render(0, 1, sampler=orc.SAMPLER_REF, threads=threads)
    dt1 = time.perf_counter() - t0
    spp = max(1, min(64, o.info.spp_used - 1, int(target_seconds / max(dt1, 1e-3))))
    t0 = time.perf_counter()
    _, st = o.render(1, 1 + spp, sampler=orc.SAMPLER_REF, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n_px * spp / dt / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port",
            "sample": f"{spp} strata x {n_px} pixels = {n_px * spp} paths of the same workload in {dt:.1f} s "
                      f"(oracle, reference sampler mode, all host threads)",
            "segments_per_path": st["segments"] / st["paths"]}


def scene_upload_bytes(built) -> int:
    d = built.desc.contents
    from surely_raytracing_b200 import capi
    n = d.n_objects * ctypes.sizeof(capi.RtbObject) + d.n_children * 4 + d.n_lights * 4
    n += d.n_materials * ctypes.sizeof(capi.RtbMaterial) + d.n_textures * ctypes.sizeof(capi.RtbTexture)
    n += d.n_perlins * ctypes.sizeof(capi.RtbPerlin)
    for i in range(d.n_images):
        n += 3 * d.images[i].width * d.images[i].height
    return n


def apply_options(scene, args):
    for o in args.option:
        k, _, v = o.partition("=")
        scene.set_option(int(k), int(v))
    return scene


def accept_image(workload, variant, mean, n_samples):
    """variance-aware acceptance of a full-size image against the committed oracle render: the full-size
    image is box-filtered 5x5 down to the oracle's resolution (same integral per block)."""
    import numpy as np
    from tests import util
    name = f"oracle_{workload}" + ("_lights" if variant else "")
    path = util.GOLDEN / f"{name}.npz"
    if not path.exists():
        return None
    gold = np.load(path)
    gh, gw = gold["mean"].shape[:2]
    H, W = mean.shape[:2]
    if (H // 5, W // 5) != (gh, gw):
        return None
    small = mean[: gh * 5, : gw * 5].reshape(gh, 5, gw, 5, 3).mean(axis=(1, 3))
    ok, rep = util.image_acceptance(small, 25 * n_samples, gold["mean"].astype(np.float64), int(gold["spp"]), gold["var"].astype(np.float64))
    img_g = np.clip(small, 0, None)
    return {"golden": f"tests/golden/{name}.npz", "accepted": bool(ok),
            "rmse": [rep[c]["rmse"] for c in range(3)], "rmse_limit": [rep[c]["lim_rmse"] for c in range(3)],
            "bias": [rep[c]["bias"] for c in range(3)], "bias_limit": [rep[c]["lim_bias"] for c in range(3)],
            "mean_radiance": float(img_g.mean())}


def time_config(cfg, pipeline, args, dev, torch, roof_peak):
    """one BASELINE config at FULL size on one GPU: device-resident and end-to-end Mpaths/s of a whole render"""
    import numpy as np
    from surely_raytracing_b200 import BuiltScene, Scene
    built = BuiltScene(cfg)
    scene = apply_options(Scene(built), args)
    info = scene.info
    H, W, n = info.image_height, info.image_width, info.spp_used
    paths = H * W * n
    accum = torch.zeros((H, W, 4), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    scene.render_device(accum.data_ptr(), 0, n, stream=stream, pipeline=pipeline)      # warm-up (workspace, instantiation)
    torch.cuda.synchronize(dev)
    ms = []
    for _ in range(3):
        accum.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scene.render_device(accum.data_ptr(), 0, n, stream=stream, pipeline=pipeline)
        e1.record()
        torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
    launches = scene.render_stats()["kernel_launches"]
    mean = scene.accum_to_pixels(accum.data_ptr()) / n
    host = np.zeros((H, W, 3))
    e2e = []
    for _ in range(3):
        host[:] = 0
        t0 = time.perf_counter()
        s2 = apply_options(Scene(built), args)
        s2.render(0, n, pipeline=pipeline, out=host)
        s2.close()
        e2e.append(time.perf_counter() - t0)
    scratch = torch.zeros((H, W, 4), dtype=torch.int64, device=dev)
    scene.render_device(scratch.data_ptr(), 0, min(n, 4), stream=stream, pipeline=pipeline, collect_stats=True)
    torch.cuda.synchronize(dev)
    st = scene.render_stats()
    i_path, _, _ = algorithmic_instr(st, info.n_lights)
    best = min(ms)
    out = {"workload": cfg, "image": [W, H], "spp": n, "max_depth": info.max_depth, "paths": paths, "ms": best,
           "value": paths / best / 1e3, "e2e": paths / min(e2e) / 1e6, "unit": "Mpaths/s", "gpu_launches": int(launches),
           "segments_per_path": st["segments"] / st["paths"], "i_path_lane_instr": i_path,
           "roofline_frac": paths / (best * 1e-3) * i_path / 1e12 / roof_peak,
           "image_check": accept_image(cfg, 0, mean, n)}
    scene.close()
    return out


def run_curve(args, dev, torch):
    """BASELINE config 5 "variance-vs-time": RMSE of the GPU image against the committed oracle render at growing spp,
    with the wall time of the call (scene upload + render + read-back), light-list mixture sampling (lights = [quad,
    sphere], reference src/main.rs:485-511) against material-only sampling (lights = empty, what render_par passes)."""
    import numpy as np
    from surely_raytracing_b200 import BuiltScene, Scene, capi
    from tests import util
    gold = np.load(util.GOLDEN / "oracle_c5.npz")
    gw = int(gold["width"])
    ref = gold["mean"].astype(np.float64)
    rows = []
    for label, drop_lights in (("mixture pdf: lights = [quad, sphere]", False), ("material pdf only: lights = []", True)):
        for spp in (1, 4, 16, 64, 256, 961, 3969, 16384):
            built = BuiltScene("c5", width=gw, spp=spp)
            if drop_lights:
                built.desc.contents.n_lights = 0
            Scene(built).render()                                   # warm-up of this size (buffers, instantiation)
            t0 = time.perf_counter()
            s = Scene(built)
            px, st = s.render()
            wall = time.perf_counter() - t0
            n = s.info.spp_used
            s.close()
            mean = np.clip(px / n, 0, 10)
            rmse = float(np.sqrt(((mean - np.clip(ref, 0, 10)) ** 2).mean()))
            rows.append({"sampling": label, "spp": n, "paths": st["paths"], "wall_ms": wall * 1e3, "device_ms": st["device_ms"],
                         "rmse_vs_oracle": rmse, "mean": float(mean.mean())})
            print(json.dumps(rows[-1]), file=sys.stderr)
    noise_floor = float(np.sqrt(np.clip(gold["var"], 0, None).mean() / int(gold["spp"])))
    print(json.dumps({"metric": "variance-vs-time (BASELINE config 5)", "workload": f"c5 at {gw}x{gw}", "oracle_spp": int(gold["spp"]),
                      "oracle_noise_floor_rmse": noise_floor, "curve": rows}), flush=True)


def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from surely_raytracing_b200 import BuiltScene, Scene, capi
    from surely_raytracing_b200.distributed import pass_rows, reduce_to_root, strong_pass

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pipeline = {"default": capi.PIPELINE_DEFAULT, "mega": capi.PIPELINE_MEGAKERNEL, "wavefront": capi.PIPELINE_WAVEFRONT}[args.pipeline]
    if args.curve:
        if rank == 0:
            run_curve(args, dev, torch)
        return

    built = BuiltScene(args.workload, width=args.width, flags=args.flags)
    t0 = time.perf_counter()
    scene = apply_options(Scene(built, device=local_rank), args)
    upload_ms = (time.perf_counter() - t0) * 1e3
    info = scene.info
    W, H, sq = info.image_width, info.image_height, info.sqrt_spp
    n_px = W * H
    R = max(1, min(args.rows_per_step, sq))
    strong = args.scaling == "strong"
    paths_per_step = n_px * sq * R * (1 if strong else world)           # whole job, all ranks

    def step_range(k):
        """stratum range this rank renders in pass k"""
        return strong_pass(k, world, rank, sq, R) if strong else pass_rows(k, world, rank, sq, R)

    accum = torch.zeros((H, W, 4), dtype=torch.int64, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up -------------------------------------------------------------------------------
    for k in range(args.warmup):
        lo, hi = step_range(k)
        scene.render_device(accum.data_ptr(), lo, hi, stream=stream, pipeline=pipeline)
    torch.cuda.synchronize(dev)
    accum.zero_()

    # ---- device-resident timing: K steps, CUDA events per step, L2 flushed between steps ------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                           # L2 flush, outside the step's events
        lo, hi = step_range(args.warmup + k)
        starts[k].record()
        scene.render_device(accum.data_ptr(), lo, hi, stream=stream, pipeline=pipeline)
        stops[k].record()
    starts[-1].record()
    reduce_to_root(accum, 0)                                    # the one collective of the job (int64 sums: exact)
    stops[-1].record()
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    step_ms = [starts[k].elapsed_time(stops[k]) for k in range(args.steps)]
    reduce_ms = starts[-1].elapsed_time(stops[-1]) if world > 1 else 0.0
    if rank == 0:
        print("step_ms", [round(x, 2) for x in step_ms], file=sys.stderr)
    timed_ms = sum(step_ms) + reduce_ms
    t = torch.tensor([timed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    timed_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = scene.render_stats()["kernel_launches"]
    value = paths_per_step * args.steps / (timed_ms * 1e-3) / 1e6

    # sanity + acceptance of the reduced image (inside the run, after the timed region)
    image_check = None
    if rank == 0:
        count = accum[..., 3]
        total_strata = args.steps * sq * R * (1 if strong else world)
        assert int(count.min()) == int(count.max()) == total_strata, (int(count.min()), int(count.max()), total_strata)
        n_grids = total_strata / (sq * sq)
        if total_strata % (sq * sq) == 0:      # whole stratum grids: the image is a complete render (n_grids times over)
            mean = scene.accum_to_pixels(accum.data_ptr()) / total_strata
            assert np.isfinite(mean).all()
            image_check = accept_image(args.workload, 0, mean, sq * sq) if not args.width else None
            if image_check is not None:
                image_check["stratum_grids_rendered"] = n_grids
                assert image_check["accepted"], image_check
        else:
            image_check = {"skipped": f"{total_strata} strata per pixel is no whole number of {sq}x{sq} stratum grids "
                                      "(a partial grid covers only a band of each pixel)"}

    # ---- e2e: through the public calls with HOST buffers.  Every step: scene upload (flatten + BVH build + H2D),
    #      render, [N > 1: int64 NCCL sum-reduce of the step's device buffer onto rank 0], ONE read-back to the host image
    h2d = scene_upload_bytes(built)
    d2h = n_px * 24
    host_px = np.zeros((H, W, 3), dtype=np.float64)
    e2e_steps = max(1, min(args.steps, 5))
    step_buf = torch.zeros((H, W, 4), dtype=torch.int64, device=dev) if world > 1 else None

    def e2e_step(k):
        lo, hi = step_range(k)
        t_a = time.perf_counter()
        s2 = apply_options(Scene(built, device=local_rank), args)   # flatten + BVH build + H2D of the scene
        t_b = time.perf_counter()
        if world == 1:
            s2.render(lo, hi, pipeline=pipeline, out=host_px)       # kernels + D2H of the sums + f64 accumulate
        else:
            step_buf.zero_()
            s2.render_device(step_buf.data_ptr(), lo, hi, stream=stream, pipeline=pipeline)
            dist.reduce(step_buf, dst=0, op=dist.ReduceOp.SUM)      # device buffers, int64: exact and split-invariant
            if rank == 0:
                s2.accum_to_pixels(step_buf.data_ptr(), out=host_px)
            else:
                torch.cuda.synchronize(dev)
        t_c = time.perf_counter()
        s2.close()
        return (round((t_b - t_a) * 1e3, 2), round((t_c - t_b) * 1e3, 2), round((time.perf_counter() - t_c) * 1e3, 2))

    for k in range(2):                                           # untimed warm-up of this path (cached buffers of the call's size)
        e2e_step(k)
    host_px[:] = 0.0
    barrier()
    e0 = time.perf_counter()
    e2e_ms = [e2e_step(args.warmup + k) for k in range(e2e_steps)]
    barrier()
    e2e_s = time.perf_counter() - e0
    if rank == 0:
        print("e2e (create, render [+ reduce] + read-back, destroy) ms", e2e_ms, file=sys.stderr)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step * e2e_steps / float(te.item()) / 1e6

    # ---- roofline: frozen algorithmic constants x the traversal's own counters of a short counted pass -----------
    roof = None
    cpu = None
    configs = None
    if rank == 0:
        lo, hi = pass_rows(0, 1, 0, sq)
        scratch = torch.zeros((H, W, 4), dtype=torch.int64, device=dev)
        scene.render_device(scratch.data_ptr(), lo, lo + 4, stream=stream, pipeline=pipeline, collect_stats=True)
        torch.cuda.synchronize(dev)
        st = scene.render_stats()
        i_path, i_extend, i_shade = algorithmic_instr(st, info.n_lights)
        peaks, how = measured_peaks()
        props = torch.cuda.get_device_properties(dev)
        n_sm = props.multi_processor_count
        sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        peak = n_sm * PEAK_LANE_INSTR_PER_CLK_PER_SM * sm_mhz * 1e6 / 1e12          # Tlane-instr/s
        per_gpu_paths_step = n_px * (step_range(0)[1] - step_range(0)[0])
        step_s = sum(step_ms) / len(step_ms) * 1e-3
        per_gpu_paths_s = per_gpu_paths_step / step_s
        achieved = per_gpu_paths_s * i_path / 1e12
        seg_per_path = st["segments"] / st["paths"]
        # per-kernel: one profiled pass (per-stage CUDA events; it synchronises every iteration, so only the SHARES are used)
        kernels = None
        if pipeline != capi.PIPELINE_MEGAKERNEL:
            scene.set_option(capi.OPT_PROFILE, 1)
            lo2, hi2 = step_range(args.warmup)
            scene.render_device(scratch.data_ptr(), lo2, hi2, stream=stream, pipeline=pipeline)
            torch.cuda.synchronize(dev)
            sp = scene.render_stats()
            scene.set_option(capi.OPT_PROFILE, 0)
            g_ms, e_ms, s_ms = sp["stage_ms"]
            tot = max(g_ms + e_ms + s_ms, 1e-9)
            segs_step = per_gpu_paths_step * seg_per_path
            kernels = {}
            for name, ms_k, i_k in (("k_wf_extend", e_ms, i_extend), ("k_wf_shade", s_ms, i_shade)):
                t_k = step_s * ms_k / tot                                     # this kernel's share of the timed step
                kernels[name] = {"share_of_step": ms_k / tot, "algorithmic_lane_instr_per_segment": i_k,
                                 "frac": segs_step * i_k / t_k / 1e12 / peak,
                                 "executed_thread_instr_per_segment": NCU["thread_instr_per_segment"][name],
                                 "useful_instruction_fraction": i_k / NCU["thread_instr_per_segment"][name],
                                 "executed_lane_frac": NCU["lanes"][name] / 32.0}
            kernels["k_wf_generate"] = {"share_of_step": g_ms / tot}
        traffic = None
        if args.workload == WORKLOAD and pipeline != capi.PIPELINE_MEGAKERNEL:   # bytes per step, like `achieved`
            traffic = NCU["dram_bytes_per_segment_c4"] * seg_per_path * per_gpu_paths_step
        executed = sum(NCU["thread_instr_per_segment"].values()) * seg_per_path
        roof = {"bound": "fp32_issue", "achieved": achieved, "peak": peak, "unit": "Tlane-instr/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": f"ncu dram bytes per segment ({NCU['source']}) x segments per step",
                "peak_source": f"{how}: {n_sm} SMs x 128 lanes x {sm_mhz:.0f} MHz",
                "i_path_lane_instr": i_path, "i_const": I_CONST,
                "i_const_source": "SURVEY.md 8(a) per-call estimates, frozen (not derived from this implementation's SASS)",
                "useful_instruction_fraction": i_path / executed, "executed_lane_instr_per_path": executed,
                "ncu_source": NCU["source"], "kernels": kernels,
                "segments_per_path": seg_per_path,
                "node_visits_per_segment": st["node_visits"] / st["segments"],
                "prim_tests_per_segment": st["prim_tests"] / st["segments"],
                "exact_tests_per_segment": st["exact_tests"] / st["segments"],
                "overflow_rays_per_segment": st["overflow_rays"] / st["segments"],
                "medium_probes_per_segment": st["medium_probes"] / st["segments"],
                "kernel": "k_render_mega" if pipeline == capi.PIPELINE_MEGAKERNEL else "k_wf_extend + k_wf_shade (whole step)",
                "hbm_secondary": {"algorithmic_gbs": per_gpu_paths_s * seg_per_path * 176 / 1e9,
                                  "measured_gbs": (traffic / step_s / 1e9) if traffic else None,
                                  "peak_gbs": float(peaks.get("hbm_gbs", 6650.0)), "bytes_per_segment": 176}}
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args.workload, args.width, args.cpu_seconds)
        if world == 1 and not args.no_configs and args.workload == WORKLOAD and not args.width:
            configs = [time_config(c, pipeline, args, dev, torch, peak) for c in ("c1", "c2", "c3", "c5")]

    if rank == 0:
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": timed_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64 geometry decisions / f32 cull + shading / i64 fixed-point accumulation", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC if args.workload == WORKLOAD else args.workload, "image": [W, H],
                       "max_depth": info.max_depth, "paths_per_step": paths_per_step, "rows_per_step": R,
                       "pipeline": args.pipeline, "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset)",
                       "scene_upload_ms": upload_ms, "surface_prims": info.n_surface_prims,
                       "bvh_nodes": info.n_bvh_nodes, "reduce_ms": reduce_ms, "wall_ms_timed_region": wall_ms,
                       "options": args.option},
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks, "roofline": roof, "image_check": image_check,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
