#!/usr/bin/env python
"""SASS evidence for the solver kernels: opcode histogram of each kernel of libgync_b200.so (cuobjdump -sass), the
tensor-memory / local-memory / FP64 counts the review asks for, and the same counts inside the rolled stage loop of the
Rosenbrock step (the largest backward branch target range that contains LDTM).
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gync.jl_b200", "csrc", "libgync_b200.so")
KERNELS = ("gync_loglik_tm_kernel", "gync_solve_tm_kernel", "gync_mcmc_persist_kernel")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
print(f"# {os.path.relpath(LIB, ROOT)}   cubin arch: {', '.join(arch)}")
cur, funcs = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2)))


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]


KEY = ("DFMA", "DMUL", "DADD", "MUFU", "LDTM", "STTM", "LDS", "STS", "LDL", "STL", "LDG", "STG", "UTCMMA", "BRA", "CALL")
for name, ins in funcs.items():
    if not any(k in name for k in KERNELS):
        continue
    h = collections.Counter(opcode(t) for _, t in ins)
    print(f"\n## {name}\n   instructions: {len(ins)}   code bytes: {16 * len(ins)}")
    print("   " + "  ".join(f"{k}={h.get(k, 0)}" for k in KEY))
    print("   top opcodes: " + ", ".join(f"{k} {v}" for k, v in h.most_common(14)))
    loops = []
    for a, t in ins:
        if opcode(t) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
    best = None
    for lo, hi in loops:
        body = [t for a, t in ins if lo <= a <= hi]
        if any(opcode(t) == "LDTM" for t in body) and 500 < len(body) < 4000:
            if best is None or len(body) > len(best[2]):
                best = (lo, hi, body)
    if best:
        hb = collections.Counter(opcode(t) for t in best[2])
        print(f"   rolled stage loop {best[0]:#x}..{best[1]:#x}: {len(best[2])} instructions ({16 * len(best[2])} bytes; L1.5 I-cache 32 KB)")
        print("      " + "  ".join(f"{k}={hb.get(k, 0)}" for k in KEY))
    inner = [(lo, hi) for lo, hi in loops if hi - lo > 16 * 8000]
    if inner:
        lo, hi = min(inner, key=lambda z: z[1] - z[0])
        body = [t for a, t in ins if lo <= a <= hi]
        hb = collections.Counter(opcode(t) for t in body)
        print(f"   persistent-lane loop (one step attempt per trip) {lo:#x}..{hi:#x}: {len(body)} instructions")
        print("      " + "  ".join(f"{k}={hb.get(k, 0)}" for k in KEY))
