#!/usr/bin/env python
"""bench.py -- Mpaths/s (camera samples/s) of the hot path on BASELINE.json's headline config:
the book-2 final scene (c4) at 800x800, depth 40, on N B200s of one box.

A "step" is one progressive pass per GPU: one rtb_render call over 10 rows of the 100x100 stratum grid
(1000 strata for every one of the 640,000 pixels = 640 M paths per GPU per step; --rows-per-step).
Row blocks are dealt round-robin to the ranks
(weak scaling: per-GPU work is fixed), each rank accumulates into its own fp32 buffer and ONE NCCL
sum-reduce at the end of the timed region delivers the image to rank 0.

Printed: ONE JSON line (see the task contract) with `value` (device-resident), `e2e` (through the
host-buffer C-ABI call: scene upload + render + read-back every step), `roofline`, `cpu_baseline`,
`clocks`, `gpu_launches`.  `--impl reference` times the CPU restatement of the reference (the
oracle: the Rust reference cannot be built in this image) on the host cores instead.
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
                 "(what main.rs passes); step = 10 rows (1000 strata) of the 100x100 stratum grid per GPU")
L2_FLUSH_BYTES = 256 << 20
ROWS_PER_STEP = 10

# Algorithmic lane-instruction constants per call (fp32-issue-slot equivalents; DFMA/DADD/DMUL = 2
# slots on B200, whose FP64 pipe issues at half the FP32 rate).  Derived from the SASS of
# librtb200.so (see DESIGN.md "Roofline"); frozen here so the bench JSON is self-describing.
I_CONST = {"node_visit": 78.0, "prim_test": 150.0, "medium_probe": 45.0, "segment_shade": 260.0, "path_setup": 140.0}
PEAK_LANE_INSTR_PER_CLK_PER_SM = 128
# ncu-measured DRAM traffic of the wavefront pipeline on c4 (dram__bytes_read.sum + dram__bytes_write.sum over every
# k_wf_* launch of profiles/r01b_wavefront_launches.csv.gz: 4140.4 GB over 7.004 steps of 2.8957 G segments each)
MEASURED_DRAM_BYTES_PER_SEGMENT_C4 = 204.1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pipeline", default="default", choices=["default", "mega", "wavefront"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--width", type=int, default=0, help="override the image width (parity/debug only)")
    ap.add_argument("--rows-per-step", type=int, default=ROWS_PER_STEP,
                    help="rows of the sqrt x sqrt stratum grid rendered per step and GPU (one rtb_render call)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
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
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC if args.workload == WORKLOAD else args.workload, "sample": sample,
                   "note": "CPU restatement (oracle port) of the Rust reference: cargo/rustc absent in this image"},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "segments_per_path": segs / max(1, paths), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def cpu_baseline(args, target_seconds):
    from oracle import orc
    from surely_raytracing_b200.scenes import BuiltScene
    b = BuiltScene(args.workload, width=args.width)
    o = orc.OracleScene(b, use_bvh=True)
    n_px = o.info.image_width * o.info.image_height
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    o.render(0, 1, sampler=orc.SAMPLER_REF, threads=threads)
    dt1 = time.perf_counter() - t0
    spp = max(1, min(64, int(target_seconds / max(dt1, 1e-3))))
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


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from surely_raytracing_b200 import BuiltScene, Scene, capi
    from surely_raytracing_b200.distributed import pass_rows, reduce_to_root

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pipeline = {"default": capi.PIPELINE_DEFAULT, "mega": capi.PIPELINE_MEGAKERNEL, "wavefront": capi.PIPELINE_WAVEFRONT}[args.pipeline]

    built = BuiltScene(args.workload, width=args.width)
    t0 = time.perf_counter()
    scene = Scene(built, device=local_rank)
    upload_ms = (time.perf_counter() - t0) * 1e3
    info = scene.info
    W, H, sq = info.image_width, info.image_height, info.sqrt_spp
    n_px = W * H
    R = max(1, min(args.rows_per_step, sq))
    paths_per_step_per_gpu = n_px * sq * R

    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up -------------------------------------------------------------------------------
    for k in range(args.warmup):
        lo, hi = pass_rows(k, world, rank, sq, R)
        scene.render_device(accum.data_ptr(), lo, hi, stream=stream, pipeline=pipeline)
    torch.cuda.synchronize(dev)
    accum.zero_()

    # ---- device-resident timing: K steps, CUDA events per step, L2 flushed between steps ------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches = 0
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                           # L2 flush, outside the step's events
        lo, hi = pass_rows(args.warmup + k, world, rank, sq, R)
        starts[k].record()
        scene.render_device(accum.data_ptr(), lo, hi, stream=stream, pipeline=pipeline)
        stops[k].record()
        launches += 1
    starts[-1].record()
    reduce_to_root(accum, 0)                                    # the one collective of the job
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
    total_paths = paths_per_step_per_gpu * args.steps * world
    value = total_paths / (timed_ms * 1e-3) / 1e6

    # sanity: the reduced image must hold exactly steps*world*sq samples per pixel
    if rank == 0:
        w_min, w_max = float(accum[..., 3].min()), float(accum[..., 3].max())
        assert w_min == w_max == float(args.steps * world * sq * R), (w_min, w_max)
        assert bool(torch.isfinite(accum).all())

    # ---- e2e: host-buffer path, every step uploads the scene and reads the image back ------------------
    import numpy as np
    h2d = scene_upload_bytes(built)
    d2h = n_px * 16
    host_px = np.zeros((H, W, 3), dtype=np.float64)
    e2e_steps = max(1, min(args.steps, 5))
    for k in range(2):                                           # untimed warm-up of the host-buffer path: a call of the
        s2 = Scene(built, device=local_rank)                     # timed size, so that the cached workspace (sized by the
        lo, hi = pass_rows(k, world, rank, sq, R)                # call) is allocated before the timed region
        s2.render(lo, hi, pipeline=pipeline, out=host_px)
        s2.close()
    host_px[:] = 0.0
    barrier()
    e0 = time.perf_counter()
    e2e_ms = []
    for k in range(e2e_steps):
        lo, hi = pass_rows(args.warmup + k, world, rank, sq, R)
        t_a = time.perf_counter()
        s2 = Scene(built, device=local_rank)                    # flatten + BVH build + H2D of the scene
        t_b = time.perf_counter()
        s2.render(lo, hi, pipeline=pipeline, out=host_px)       # kernels + D2H of the sums + f64 accumulate
        t_c = time.perf_counter()
        s2.close()
        e2e_ms.append((round((t_b - t_a) * 1e3, 2), round((t_c - t_b) * 1e3, 2), round((time.perf_counter() - t_c) * 1e3, 2)))
    if rank == 0:
        print("e2e (create, render, destroy) ms", e2e_ms, file=sys.stderr)
    if world > 1:
        tsum = torch.from_numpy(host_px).to(dev)
        dist.reduce(tsum, dst=0)
        if rank == 0:
            host_px = tsum.cpu().numpy()
    barrier()
    e2e_s = time.perf_counter() - e0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step_per_gpu * e2e_steps * world / float(te.item()) / 1e6

    # ---- roofline: algorithmic lane-instructions from device counters of a short counted pass -----------
    roof = None
    cpu = None
    if rank == 0:
        lo, hi = pass_rows(0, 1, 0, sq)
        scratch = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        scene.render_device(scratch.data_ptr(), lo, lo + 4, stream=stream, pipeline=pipeline, collect_stats=True)
        torch.cuda.synchronize(dev)
        st = scene.render_stats()
        paths = st["paths"]
        i_path = (I_CONST["path_setup"] + (st["segments"] * I_CONST["segment_shade"] + st["node_visits"] * I_CONST["node_visit"]
                  + st["prim_tests"] * I_CONST["prim_test"] + st["medium_probes"] * I_CONST["medium_probe"]) / paths)
        peaks, how = measured_peaks()
        props = torch.cuda.get_device_properties(dev)
        n_sm = props.multi_processor_count
        sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        peak = n_sm * PEAK_LANE_INSTR_PER_CLK_PER_SM * sm_mhz * 1e6 / 1e12          # Tlane-instr/s
        per_gpu_paths_s = paths_per_step_per_gpu / (sum(step_ms) / len(step_ms) * 1e-3)
        achieved = per_gpu_paths_s * i_path / 1e12
        seg_per_path = st["segments"] / paths
        traffic = None
        if args.workload == WORKLOAD and pipeline != capi.PIPELINE_MEGAKERNEL:   # bytes per step, like `achieved`
            traffic = MEASURED_DRAM_BYTES_PER_SEGMENT_C4 * seg_per_path * paths_per_step_per_gpu
        roof = {"bound": "fp32_issue", "achieved": achieved, "peak": peak, "unit": "Tlane-instr/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "ncu dram bytes per segment (profiles/r01b_wavefront_launches.csv.gz) x segments per step",
                "peak_source": f"{how}: {n_sm} SMs x 128 lanes x {sm_mhz:.0f} MHz",
                "i_path_lane_instr": i_path, "segments_per_path": seg_per_path,
                "node_visits_per_segment": st["node_visits"] / st["segments"],
                "prim_tests_per_segment": st["prim_tests"] / st["segments"],
                "medium_probes_per_segment": st["medium_probes"] / st["segments"],
                "kernel": "k_render_mega" if pipeline == capi.PIPELINE_MEGAKERNEL else "k_wf_extend (+ k_wf_shade)",
                "hbm_secondary": {"algorithmic_gbs": per_gpu_paths_s * seg_per_path * 176 / 1e9,
                                  "measured_gbs": (traffic / (sum(step_ms) / len(step_ms) * 1e-3) / 1e9) if traffic else None,
                                  "peak_gbs": float(peaks.get("hbm_gbs", 6650.0)), "bytes_per_segment": 176}}
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args, args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": timed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 geometry / f32 shading", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC if args.workload == WORKLOAD else args.workload, "image": [W, H],
                       "max_depth": info.max_depth, "paths_per_step_per_gpu": paths_per_step_per_gpu, "rows_per_step": R,
                       "pipeline": args.pipeline, "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset)",
                       "scene_upload_ms": upload_ms, "surface_prims": info.n_surface_prims,
                       "bvh_nodes": info.n_bvh_nodes, "reduce_ms": reduce_ms, "wall_ms_timed_region": wall_ms},
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks, "roofline": roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
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
