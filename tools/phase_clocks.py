#!/usr/bin/env python
"""Where a small-frame CTA spends its time: cycles per phase of pc_scan_small_kernel.

    python tools/phase_clocks.py --build          # CPU box: profiling build (-DPC_PHASE_CLOCKS)
    PYCONTACT_B200_LIB=pycontact_b200/libpycontact_b200_prof.so python tools/phase_clocks.py [workload] [frames]

Thread 0 of every CTA reads clock64() after each CTA barrier; the differences are summed over CTAs.
Not a bench number (the marks cost a few atomics per frame); the SHARES are what is read.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PROF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pycontact_b200", "libpycontact_b200_prof.so")

if "--build" in sys.argv:
    from pycontact_b200 import build
    print(build.build_variant(PROF, ["PC_PHASE_CLOCKS"]))
    sys.exit(0)

import bench
from pycontact_b200 import _lib, device_synth
from pycontact_b200._lib import DeviceBuffer, check
from pycontact_b200.engine import ContactEngine

NAMES = ["1 bbox pass (HBM read)", "2 grid (warp 0)", "3 stage in-grid atoms (L2 read)", "4 scans", "5 scatter",
         "6 search (thread 0)", "6 search: wait for the slowest warp + table init", "7 claim output range",
         "7a drain: records + accumulation (thread 0)", "7a wait for the slowest warp", "8 rows / H-bonds -> global",
         "7b H-bond tests on the gated pairs"]

args = [x for x in sys.argv[1:] if not x.startswith("--")]
name = args[0] if args else "c2"
F = int(args[1]) if len(args) > 1 else 2048
w = bench.prepare_workload(name)
s, s1, s2, gm1, gm2 = w["system"], w["s1"], w["s2"], w["gm1"], w["gm2"]
lib = _lib.load()
if not hasattr(lib, "pc_phase_clocks"):
    raise SystemExit("set PYCONTACT_B200_LIB to the profiling build (tools/phase_clocks.py --build)")
n1, n2 = s1.n, s2.n
gen1 = device_synth.DeviceTrajectory(s, 0, s.sel1); d1 = DeviceBuffer(F * n1 * 12, 0); gen1.generate(d1.ptr, 0, F)
gen2 = device_synth.DeviceTrajectory(s, 0, s.sel2); d2 = DeviceBuffer(F * n2 * 12, 0); gen2.generate(d2.ptr, 0, F)
check(lib.pc_device_sync(0))
lib.pc_phase_clocks.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]
for ctas in [int(x) for x in (args[2].split(',') if len(args) > 2 else ['0', '1'])]:
    eng = ContactEngine(s1, s2, 5.0, 2.5, 120, device=0, slots=1)
    eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
    lib.pc_set_small_ctas(eng.slots[0].ctx, ctas)
    for rep in range(3):
        tot = eng.scan_device(d1.ptr, d2.ptr, F, accumulate=True)
    out = (ctypes.c_ulonglong * 24)()
    lib.pc_phase_clocks(eng.slots[0].ctx, out)
    v = list(out)[:12]
    t = float(sum(v)) or 1.0
    print("workload %s, %d frames, small_ctas=%d (ran %d CTAs/SM, %d frames retried by the full plan), totals %s" % (name, F, ctas, out[15], out[14], tot))
    for nm, x in zip(NAMES, v):
        print("  %-52s %12d cycles  %5.1f%%  %8.0f cycles/frame" % (nm, x, 100.0 * x / t, x / F))
    print("  %-52s %12d cycles          %8.0f cycles/frame" % ("sum", sum(v), sum(v) / F))
    print("  frames handed to the full plan: %d for staging (in-grid atoms), %d for the hit queue" % (out[12], out[13]))
    print("  atoms inside the grid per frame: mean %.0f, max %d (%d frame passes)" % (out[16] / max(out[18], 1), out[17], out[18]))
    eng.close()
