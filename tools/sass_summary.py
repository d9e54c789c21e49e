#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (no GPU needed): evidence of what the kernels are made of.

    python tools/sass_summary.py [--so lib.so] > profiles/r2_sass_summary.txt

UBLKCP = bulk async copy (cp.async.bulk, the 1-D TMA path of pc_stream.cuh), SYNCS = mbarrier operations,
FADD2 / FMUL2 = packed float32 pipe (pair_within_packed), ATOMS / RED / ATOMG = shared / global atomics,
LDL / STL = local memory (spills), DADD / DMUL / DFMA = float64 (distance, H-bond geometry), MUFU = special function unit.
No UTC*MMA / TMEM instructions are expected: a cutoff search is not a contraction.
"""
import argparse, collections, os, re, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pycontact_b200", "libpycontact_b200.so"))
a = ap.parse_args()
txt = subprocess.run(["cuobjdump", "-sass", a.so], capture_output=True, text=True).stdout
keys = ["UBLKCP", "SYNCS", "FADD2", "FMUL2", "FFMA", "LDS", "STS", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDL", "STL", "DADD", "DMUL", "DFMA", "MUFU", "SHFL", "VOTE", "BAR", "UTC", "HMMA"]
per = collections.OrderedDict()
cur = None
arch = None
for ln in txt.split("\n"):
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = per.setdefault(name[-60:], collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for k in keys:
            if op.startswith(k):
                cur[k] += 1
print("library:", os.path.basename(a.so), "arch:", arch)
print("%-60s %6s " % ("kernel", "instr") + " ".join("%6s" % k for k in keys))
for name, c in per.items():
    print("%-60s %6d " % (name, c["total"]) + " ".join("%6d" % c[k] for k in keys))
