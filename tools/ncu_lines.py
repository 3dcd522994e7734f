#!/usr/bin/env python
"""Join an ncu source-page export with nvdisasm line info and aggregate per source function / line.

usage: tools/ncu_lines.py <report.ncu-rep> <library.so> <mangled-kernel-substring> <demangled-substring> [--lines N]
Prints, per device function of csrc/rtb_device.cuh / kernels.cu: warp-instructions executed,
thread-instructions executed, average active lanes, and stall samples.  Used to write profiles/*.md."""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path


def line_map(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(so).resolve())], cwd=tmp, capture_output=True)
    out = {}
    for cubin in Path(tmp).glob("*.cubin"):
        txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
        cur_fn, cur = None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (Path(m.group(1)).name, int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and kernel_sub in cur_fn:
                out.setdefault(cur_fn, {})[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def function_ranges(path):
    """(start_line, name) of every function definition in a source file (crude but sufficient)."""
    res = []
    for i, ln in enumerate(Path(path).read_text().splitlines(), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:RTB_DEV|__global__|__device__|static|inline|cudaError_t|size_t)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", ln)
        if m and not ln.strip().startswith("//") and m.group(1) not in ("defined", "if", "for", "while", "return"):
            res.append((i, m.group(1)))
    return res


def main():
    rep, so, ksub, dsub = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = raw.split('"Kernel Name",')
    root = Path(__file__).resolve().parent.parent / "surely_raytracing_b200" / "csrc"
    ranges = {p.name: function_ranges(p) for p in [root / "rtb_device.cuh", root / "kernels.cu", root / "wavefront.cu"]}
    lm_all = line_map(so, ksub)
    for blk in blocks[1:]:
        kname, rest = blk.split("\n", 1)
        if dsub not in kname:
            continue
        rows = list(csv.reader(io.StringIO(rest)))
        hdr = rows[0]
        ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        data = [(int(r[ia], 16), float(r[ii] or 0), float(r[it] or 0), float(r[isamp] or 0), r[1]) for r in rows[1:] if len(r) > it and r[ia].startswith("0x")]
        base = min(d[0] for d in data)
        # choose the line map whose instruction count matches
        lm = None
        for fn, m in lm_all.items():
            if abs(len(m) - len(data)) <= 2:
                lm = m
        if lm is None:
            lm = max(lm_all.values(), key=len)
        per_fn = defaultdict(lambda: [0.0, 0.0, 0.0])
        per_line = defaultdict(lambda: [0.0, 0.0, 0.0])
        tot = [0.0, 0.0, 0.0]
        for addr, wi, ti, smp, sass in data:
            (f, l), _ = lm.get(addr - base, (("?", 0), ""))
            name = "?"
            for s, n in ranges.get(f, []):
                if s <= l:
                    name = n
            key = f"{f}:{name}"
            for acc in (per_fn[key], per_line[(f, l)], tot):
                acc[0] += wi; acc[1] += ti; acc[2] += smp
        print(f"kernel {kname.strip().strip(',').strip(chr(34))[:90]}")
        print(f"total warp-instr {tot[0]:.4g}  thread-instr {tot[1]:.4g}  avg lanes {tot[1] / max(tot[0], 1):.2f}  samples {tot[2]:.0f}")
        print(f"{'function':48s} {'warp-instr%':>11s} {'thread-instr%':>13s} {'lanes':>6s} {'stall-samples%':>14s}")
        for k, v in sorted(per_fn.items(), key=lambda kv: -kv[1][0]):
            print(f"{k:48s} {100 * v[0] / tot[0]:11.2f} {100 * v[1] / tot[1]:13.2f} {v[1] / max(v[0], 1):6.2f} {100 * v[2] / max(tot[2], 1):14.2f}")
        print(f"\ntop {nlines} source lines by warp-instructions")
        for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:nlines]:
            print(f"{f}:{l:<5d} {100 * v[0] / tot[0]:6.2f}% warp-instr  lanes {v[1] / max(v[0], 1):5.2f}  samples {100 * v[2] / max(tot[2], 1):5.2f}%")


if __name__ == "__main__":
    main()
