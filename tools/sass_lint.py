#!/usr/bin/env python
"""SASS lint for the ptxas hazard that produced round 1's "illegal memory access" in the fused trace kernel.

What happened there (cuda-gdb, profiles/r02_o3_fault_cuda_gdb.txt): ptxas 12.9 hoisted the
loop-invariant address `prims + 0x30` into the UNIFORM register pair UR5:UR6 before the divergent BVH traversal loop
(a BSSY.RELIABLE region that lanes leave one by one through `BREAK.RELIABLE` + a branch past the BSYNC), and re-used UR5
for an f64 immediate (`UMOV UR5, 0x3fd45f30`, a coefficient of the atan2 polynomial) in the code after the loop.  Uniform
registers are per WARP: the lanes that left the loop early overwrote UR5 while the others were still traversing with it
-- their next primitive load went to {0x3fd45f30 + 96 * prim, hi} = a wild address.  Nothing in the source is undefined;
-O1 / assert builds merely allocate differently.

The lint flags every kernel in which a uniform register is (a) carried INTO A LOOP that lies inside a BSSY.RELIABLE
region with a BREAK.RELIABLE exit (read in the loop before any write in the loop: a hoisted loop invariant that must
survive while other lanes of the warp run elsewhere) and (b) written again at a higher address after the region.
(Uniform immediates re-materialised right before their use -- UMOV; use -- are everywhere and harmless.)
tests/test_abi.py runs it over the shipped library: no kernel may show the pattern.

usage: tools/sass_lint.py <library.so> [kernel-substring]"""
import re
import subprocess
import sys

INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
UR = re.compile(r"\bUR(\d+)\b")


def kernels(so):
    """(name, [(address, instruction text)]) per kernel of a library, or of a saved SASS listing (.txt)"""
    txt = open(so).read() if so.endswith(".txt") else subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    name, body = None, []
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = INSTR.match(ln)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


def ur_defs_uses(text):
    """(written URs, read URs) of one SASS instruction"""
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    op, _, rest = t.partition(" ")
    ops = [o.strip() for o in rest.split(",")]
    wide = ".64" in op or "WIDE" in op
    writes, reads = set(), set()
    uniform_dst = op.startswith(("UMOV", "UIADD", "ULDC", "LDCU", "S2UR", "ULOP", "USHF", "UIMAD", "ULEA", "USEL", "UPRMT", "UFLO", "UPOPC",
                                 "UBREV", "R2UR", "VOTEU", "UF2", "UI2", "UCGABAR", "REDUX", "UMEMSETS"))
    start = 0
    if uniform_dst and ops:
        for m in UR.finditer(ops[0]):
            n = int(m.group(1))
            writes.add(n)
            if wide or op.startswith("LDCU.64"):
                writes.add(n + 1)
        start = 1
        if op.startswith("UIADD3") and len(ops) > 1 and ops[1].startswith("UP"):
            start = 2
    for o in ops[start:]:
        for m in UR.finditer(o):
            n = int(m.group(1))
            reads.add(n)
            if "desc[" in o or wide or ".64" in o:
                reads.add(n + 1)
    return writes, reads


def lint(name, body):
    """findings: (region start, region end, UR, address of the later write, its text)"""
    findings = []
    for i, (a, t) in enumerate(body):
        m = re.search(r"BSSY\.RELIABLE\s+B\d+,\s*0x([0-9a-f]+)", t)
        if not m:
            continue
        end = int(m.group(1), 16)
        j_end = max((k for k, (b, _) in enumerate(body) if b <= end), default=i)
        region = body[i:j_end + 1]
        if not any("BREAK.RELIABLE" in x for _, x in region):
            continue
        # loops (backward branches) that lie inside the region or wrap it (the loop re-arms the barrier every trip)
        hoisted = set()
        for b, x in body:
            mb = re.search(r"\bBRA\b.*\b0x([0-9a-f]+)\s*$", x)
            if not mb:
                continue
            tgt = int(mb.group(1), 16)
            if not (tgt < b and tgt <= end and b >= a):
                continue
            loop = [(c, y) for c, y in body if tgt <= c <= b]
            defined = set()
            for _, y in loop:   # read in the loop before any write in the loop = a value carried in from outside it
                w, r = ur_defs_uses(y)
                hoisted |= {u for u in r if u not in defined}
                defined |= w
        loop_end = max([b for b, x in body if re.search(r"\bBRA\b.*\b0x([0-9a-f]+)\s*$", x) and
                        int(re.search(r"0x([0-9a-f]+)\s*$", x).group(1), 16) < b and int(re.search(r"0x([0-9a-f]+)\s*$", x).group(1), 16) <= end and b >= a] + [end])
        # lanes ESCAPE only if a BREAK.RELIABLE is followed by a branch to beyond the loop (past its BSYNC and back-edge);
        # a break that lands on a barrier inside the loop reconverges before anything is rewritten
        escapes = False
        for k, (b, x) in enumerate(region):
            if "BREAK.RELIABLE" in x:
                for b2, x2 in region[k + 1:k + 3]:
                    mt = re.search(r"\bBRA\b.*\b0x([0-9a-f]+)\s*$", x2)
                    if mt and int(mt.group(1), 16) > loop_end:
                        escapes = True
        if not escapes:
            continue
        # out-of-line subroutines (CALL.REL.NOINC targets: 64-bit division, f64 slow paths) sit behind the kernel's own
        # code; what they write counts where they are CALLED, not where they are stored
        subs = sorted({int(m2.group(1), 16) for _, x in body for m2 in [re.search(r"\bCALL\.REL\.NOINC\s+0x([0-9a-f]+)", x)] if m2})
        first_sub = subs[0] if subs else None

        def sub_writes(target):
            out = []
            for b2, x2 in body:
                if b2 < target:
                    continue
                out.append((b2, x2))
                if x2.startswith("RET"):
                    break
            return out

        for b, x in body:
            if b <= loop_end:
                continue
            if first_sub is not None and b >= first_sub:
                break
            if re.match(r"(@!?U?P\d+\s+)?WARPSYNC(?!\.COLLECTIVE)", x):
                break   # __syncwarp of the lanes that entered the loop: nobody runs ahead past this point
            mc = re.search(r"\bCALL\.REL\.NOINC\s+0x([0-9a-f]+)", x)
            for b2, x2 in ([(b, x)] if not mc else sub_writes(int(mc.group(1), 16))):
                w, _ = ur_defs_uses(x2)
                for u in sorted(w & hoisted):
                    findings.append((a, end, u, b2, x2))
    return findings


def main():
    so = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    total = 0
    for name, body in kernels(so):
        if sub not in name:
            continue
        f = lint(name, body)
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:90]
        if f:
            total += 1
            regs = sorted({u for _, _, u, _, _ in f})
            print(f"HAZARD {short}: uniform registers {['UR%d' % u for u in regs]} are live into a BREAK-exited region and rewritten after it")
            for a, end, u, b, x in f[:4]:
                print(f"    region 0x{a:04x}-0x{end:04x}: UR{u} rewritten at 0x{b:04x}: {x}")
        else:
            print(f"ok     {short}")
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
