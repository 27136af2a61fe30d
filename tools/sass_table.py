#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass): the evidence that the hot kernels are tcgen05 / TMA code.
Usage: python tools/sass_table.py > profiles/rNN_sass_per_kernel.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shadernn_b200", "libsnn_b200.so")
COLS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "UTCATOMSWS", "ELECT", "BRA.U.ANY", "SYNCS", "HADD2.F32", "F2FP", "HMMA", "R2UR"]


def demangle(name):
    out = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    out = out.replace("void ", "").replace("snnb::", "").replace("(int)", "").replace("(bool)", "")
    out = re.sub(r">\(.*", ">", out) if ">(" in out else re.sub(r"\(.*", "", out)
    return out


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    counts[cur][c] += 1
            counts[cur]["_n"] += 1
    print("# Per-kernel SASS mnemonic counts of shadernn_b200/libsnn_b200.so (cuobjdump -sass, sm_100a), tensor-core and TMA kernels.")
    print("# UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store, UBLKCP = cp.async.bulk (plain), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
    print("# UTCATOMSWS = TMEM alloc, SYNCS = mbarrier ops, ELECT = elect.sync, HMMA = legacy mma.sync (must be 0), HADD2.F32 / F2FP = split-fp16 unpack / pack,")
    print("# R2UR = register -> uniform register moves (the tax of a single-thread issue loop), instr = all SASS instructions of the kernel.")
    print("%-64s" % "kernel" + "".join("%11s" % c for c in COLS) + "%9s" % "instr")
    for f in order:
        c = counts[f]
        if not (c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"]):
            continue
        print("%-64s" % demangle(f)[:64] + "".join("%11d" % c[k] for k in COLS) + "%9d" % c["_n"])


if __name__ == "__main__":
    main()
