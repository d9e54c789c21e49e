#!/usr/bin/env python
"""Per-source-line cost of a kernel from an ncu report (CPU box, no GPU needed).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [--so lib.so] [--top 40] [--launch 0]

ncu's CSV source page lists SASS instructions with executed counts and stall samples but no line
numbers; nvdisasm -g gives the same instruction sequence with ``//## File .. line N`` markers.
The two are joined by position (the library must be the build that was profiled).
"""
import argparse
import collections
import csv
import glob
import os
import re
import subprocess
import tempfile


def sass_lines(so, kernel, sym=None, outer=False):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    out = []
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-gi" if outer else "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur_fn, cur_line, cur_file = None, 0, ""
        for ln in txt.split("\n"):
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                continue
            # with -gi an inlined instruction carries one marker per frame, innermost first; each marker
            # overwrites the previous one, so the outermost frame (the kernel's own line) wins
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and kernel in cur_fn and (sym is None or sym in cur_fn):
                out.append((int(m.group(1), 16), cur_file, cur_line, m.group(2).strip()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "pycontact_b200", "libpycontact_b200.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--outer", action="store_true", help="attribute inlined code to the line of the kernel that called it")
    ap.add_argument("--ranges", default=None, help="sum over line ranges of the kernel's file: 152-182:bbox,233-275:stage,...")
    ap.add_argument("--sym", default=None, help="substring of the mangled name (e.g. ILi512E) to pick one instantiation")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split("\n")))
    # split per kernel launch
    launches, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            launches.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    sel = [l for l in launches if args.kernel in l["name"]]
    if not sel:
        raise SystemExit("kernel not in report: %s" % [l["name"][:60] for l in launches])
    L = sel[min(args.launch, len(sel) - 1)]
    hdr = L["hdr"]
    ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    sl = sass_lines(args.so, args.kernel.split("<")[0], args.sym, args.outer or bool(args.ranges))
    n = min(len(sl), len(L["rows"]))
    if len(sl) != len(L["rows"]):
        print("warning: %d SASS rows in report vs %d in library (joined by position)" % (len(L["rows"]), len(sl)))
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_inst = tot_samp = 0
    for k in range(n):
        r = L["rows"][k]
        _, f, line, _ = sl[k]
        inst = int(float(r[ci["Instructions Executed"]] or 0))
        samp = int(float(r[ci["# Samples"]] or 0))
        thr = int(float(r[ci["Thread Instructions Executed"]] or 0))
        a = agg[(f, line)]
        a[0] += inst; a[1] += samp; a[2] += thr
        for i in stall_cols:
            v = int(float(r[i] or 0))
            if v:
                a[3][hdr[i]] += v
        tot_inst += inst; tot_samp += samp
    print("kernel %s: %d warp-instructions, %d samples" % (L["name"][:70], tot_inst, tot_samp))
    if args.ranges:
        print("%-28s %12s %6s %8s %6s  %s" % ("range", "warp-inst", "%", "samples", "%", "top stalls"))
        rest = [0, 0, collections.Counter()]
        used = set()
        for spec in args.ranges.split(","):
            rng, name = spec.split(":")
            lo, hi = [int(x) for x in rng.split("-")]
            acc = [0, 0, collections.Counter()]
            for (f, line), a in agg.items():
                if lo <= line <= hi and f.startswith("pc_scan_small"):
                    acc[0] += a[0]; acc[1] += a[1]; acc[2].update(a[3]); used.add((f, line))
            st = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in acc[2].most_common(4))
            print("%-28s %12d %5.1f%% %8d %5.1f%%  %s" % (name, acc[0], 100.0 * acc[0] / max(tot_inst, 1), acc[1],
                                                          100.0 * acc[1] / max(tot_samp, 1), st))
        for k, a in agg.items():
            if k not in used:
                rest[0] += a[0]; rest[1] += a[1]; rest[2].update(a[3])
        print("%-28s %12d %5.1f%% %8d %5.1f%%" % ("(other)", rest[0], 100.0 * rest[0] / max(tot_inst, 1), rest[1],
                                                  100.0 * rest[1] / max(tot_samp, 1)))
        return
    print("%-24s %6s %12s %6s %8s %6s  %s" % ("file", "line", "warp-inst", "%", "samples", "%", "top stalls"))
    for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[: args.top]:
        st = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in a[3].most_common(3))
        print("%-24s %6d %12d %5.1f%% %8d %5.1f%%  %s" % (f, line, a[0], 100.0 * a[0] / max(tot_inst, 1), a[1],
                                                          100.0 * a[1] / max(tot_samp, 1), st))


if __name__ == "__main__":
    main()
