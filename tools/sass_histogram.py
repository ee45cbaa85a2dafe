#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libspatialvla_b200.so (cuobjdump -sass): which kernels carry the Blackwell tensor-core / TMEM /
TMA instructions (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce,
UTCBAR = tcgen05.commit) and which run on the warp-level MMA (HMMA) or CUDA cores.  Usage: python tools/sass_histogram.py > profiles/sass_rN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spatialvla_b200", "lib", "libspatialvla_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA", "LDSM", "LDGSTS", "DFMA", "MUFU", "FFMA",
        "ATOMG", "REDG", "REDUX"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(anonymous namespace\)::|svla_attn_tc::|^void ", "", kern).split("(")[0]
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and kern:
            hist[kern][m.group(1)] += 1
            if m.group(1) in ("UTCHMMA", "UTMALDG") and m.group(2):
                hist[kern][m.group(1) + m.group(2)] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static), {len(hist)} kernels")
    print(f"{'kernel':78s} {'total':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
    for k, c in hist.items():
        print(f"{k[:78]:78s} {sum(v for kk, v in c.items() if '.' not in kk):7d} " + " ".join(f"{c.get(x, 0):8d}" for x in KEYS))
    print("\n# variants of the tensor-core / TMA instructions")
    for k, c in hist.items():
        var = {kk: v for kk, v in c.items() if "." in kk}
        if var:
            print(f"{k[:78]:78s} " + ", ".join(f"{kk} x{v}" for kk, v in sorted(var.items())))


if __name__ == "__main__":
    sys.exit(main())
