#!/usr/bin/env python
"""SASS evidence of libmahout_b200.so: per kernel, registers and the counts of the mnemonics that show which hardware
paths are used, plus an excerpt of K3.  Runs on the build machine (cuobjdump only, no GPU):

    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mahout_b200", "libmahout_b200.so")
WATCH = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "ATOMS", "ATOMG", "RED", "MATCH", "REDUX", "UBLKPF",
         "BAR.SYNC", "IMAD.WIDE", "DFMA", "DMUL", "LDG.E.128", "LDG.E.64", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = tuple(int(x) for x in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            kernels[cur].append(line)
    names = demangle(list(kernels))
    print("SASS evidence of libmahout_b200.so (cuobjdump -sass, sm_100a), final build of round 2 (tools/sass_evidence.py).")
    print("Per kernel: registers / stack bytes / static shared memory, instruction count, and the counts of the mnemonics that")
    print("show which hardware paths are used.  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM/STTM = tcgen05.ld/st")
    print("(TMEM), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, ATOMS = shared-memory atomics, RED/ATOMG = global")
    print("reductions/atomics, MATCH = match.any, REDUX = redux.sync, UBLKPF = cp.async.bulk.prefetch.L2, STL/LDL = local")
    print("memory (spills / stack arrays).\n")
    for k in sorted(kernels, key=lambda x: names[x]):
        counts = collections.Counter()
        for line in kernels[k]:
            body = line.split("*/", 1)[1]
            for w in WATCH:
                if re.search(r"(?<![A-Z.])" + re.escape(w) + r"(?![A-Z])", body):
                    counts[w] += 1
        r = regs.get(k)
        head = f"{r[0]} registers, {r[1]} B stack, {r[2]} B static smem; " if r else ""
        print(names[k])
        print(f"    {head}{len(kernels[k])} instructions; " + ", ".join(f"{w} x{c}" for w, c in sorted(counts.items())))
    target = [k for k in kernels if "k_cosineILi256ELi1ELb1" in k]
    if target:
        print("\nExcerpt, k_cosine<256,1,true> (config 3's kernel): the TMA loads, the MMA issue, the TMEM epilogue loads / stores")
        shown = 0
        for line in kernels[target[0]]:
            if re.search(r"UTMALDG|UTCHMMA|UTCBAR|LDTM|STTM", line) and shown < 40:
                print(line.rstrip())
                shown += 1


if __name__ == "__main__":
    sys.exit(main())
