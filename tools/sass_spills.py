#!/usr/bin/env python
"""Local-memory traffic of a kernel in the built library, by source line (no GPU needed).

    python tools/sass_spills.py KERNEL_SUBSTRING [--so lib.so]

Lists every LDL / STL (register spills and stack arrays) of the kernel's SASS with the source line nvdisasm
attributes it to, and the instruction count per line for the lines that hold any -- a spill inside a hot loop
shows up as long-scoreboard stalls on lines that only touch shared memory.
"""
import argparse, collections, glob, os, re, subprocess, tempfile
ap = argparse.ArgumentParser()
ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pycontact_b200", "libpycontact_b200.so"))
a = ap.parse_args()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, check=True, capture_output=True)
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    fn, line, fil = None, 0, ""
    per = collections.defaultdict(lambda: [0, 0, 0])
    total = 0
    for ln in txt.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            fil, line = os.path.basename(m.group(1)), int(m.group(2)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and fn and a.kernel in fn:
            ins = m.group(2)
            total += 1
            k = (fn[:40], fil, line)
            per[k][0] += 1
            if re.search(r"\bLDL\b|LDL\.", ins): per[k][1] += 1
            if re.search(r"\bSTL\b|STL\.", ins): per[k][2] += 1
    print("instructions:", total)
    for k, v in sorted(per.items(), key=lambda kv: (kv[0][0], kv[0][1], kv[0][2])):
        if v[1] or v[2]:
            print("%-42s %-24s line %4d: %4d instr, %3d LDL, %3d STL" % (k[0], k[1], k[2], v[0], v[1], v[2]))
