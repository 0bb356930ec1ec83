#!/usr/bin/env python3
"""Opcode histogram of the built library's kernels (cuobjdump -sass), per kernel:
evidence of what the sm_100a code is made of (UTMALDG = 2-D tensor copies, FFMA2 / FADD2 =
packed FP32, SYNCS = mbarrier transactions, VIMNMX3 = packed 16-bit peak maxima ...).
usage: python tools/sass_histogram.py [lib] > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "loudgain_b200", "lib", "libebur128.so")
sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
kernels, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and name:
        op = m.group(1)
        base = op.split(".")[0]
        key = base
        # keep the qualifiers that say what kind of access / arithmetic it is
        for q in ("2D", "S16x2", "TRANS64", "ARRIVE", "F64", "128", "64"):
            if ("." + q) in op:
                key += "." + q
        kernels[name][key] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (static instructions), cuobjdump -sass")
for k, c in kernels.items():
    total = sum(c.values())
    print(f"\n== {k}  [{total} instructions]")
    for op, n in c.most_common(28):
        print(f"  {op:<22} {n:6d}  {100.0 * n / total:5.1f} %")
