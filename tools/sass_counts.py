#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of libpiql_b200.so the registers / shared memory (cuobjdump -res-usage) and the counts
of the instructions that show what the kernel does on sm_100a -- UTMALDG (TMA tensor box loads), UBLKCP (bulk async copies),
packed fp32 FADD2 / FMUL2 / FFMA2, MUFU.EX2 / MUFU.LG2, 256-bit LDG, 64-bit STG, SHFL, BAR, and that no tensor-core op
(UTC*MMA / HMMA / LDTM) appears: nothing on this path is a contraction.
usage: python tools/sass_counts.py [lib.so] > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "piql.jl_b200", "lib", "libpiql_b200.so")
PAT = collections.OrderedDict([
    ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("FFMA2", r"\bFFMA2"), ("FADD2", r"\bFADD2"), ("FMUL2", r"\bFMUL2"),
    ("MUFU.EX2", r"\bMUFU\.EX2"), ("MUFU.LG2", r"\bMUFU\.LG2"), ("LDG.256", r"\bLDG\.[A-Z0-9.]*256"), ("STG.64", r"\bSTG\.[A-Z0-9.]*64\b"),
    ("STG.128", r"\bSTG\.[A-Z0-9.]*128"), ("LDS.128", r"\bLDS\.128"), ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\."), ("DFMA", r"\bDFMA"),
    ("IMAD.WIDE", r"\bIMAD\.WIDE"), ("tensor", r"\b(UTC[A-Z]*MMA|HMMA|IMMA|DMMA|LDTM|UTCBAR)"),
])


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    clean = []
    for o in out:
        o = re.sub(r"^void ", "", o)
        o = (o.split(">(")[0] + ">") if ">(" in o else o.split("(")[0]      # drop the parameter list, keep the template arguments
        clean.append(o.replace("(piql::Kind)", "").replace("piql::", ""))
    return clean


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    rows = []
    for b in blocks:
        name = b.split("\n", 1)[0].strip()
        lines = [l for l in b.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        counts = [sum(1 for l in lines if re.search(p, l)) for p in PAT.values()]
        rows.append((name, len(lines), counts))
    names = demangle([r[0] for r in rows])
    print("# cuobjdump -sass / -res-usage of %s (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)" % os.path.relpath(LIB, ROOT))
    print("# columns: registers stack static-shared | instructions | " + " ".join(PAT.keys()))
    tot = [0] * len(PAT)
    for (mangled, n, counts), nm in sorted(zip(rows, names), key=lambda x: x[1]):
        reg, stack, sh = usage.get(mangled, (0, 0, 0))
        print("%-64s R%-3d S%-4d sh%-5d | %5d | %s" % (nm[:64], reg, stack, sh, n, " ".join("%d" % c for c in counts)))
        tot = [a + b for a, b in zip(tot, counts)]
    print("# totals: " + " ".join("%s=%d" % (k, v) for k, v in zip(PAT.keys(), tot)))
    assert tot[-1] == 0, "tensor-core instructions found"


if __name__ == "__main__":
    main()
