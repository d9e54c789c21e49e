"""Host-side timing of each library call for one batch (debug aid)."""
import ctypes, sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from pycontact_b200 import _lib, device_synth
from pycontact_b200._lib import DeviceBuffer, check
from pycontact_b200.engine import ContactEngine

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 22
w = bench.prepare_workload(name)
s, s1, s2, gm1, gm2 = w["system"], w["s1"], w["s2"], w["gm1"], w["gm2"]
lib = _lib.load()
n1, n2 = s1.n, s2.n if s2 is not None else s1.n
gen1 = device_synth.DeviceTrajectory(s, 0, s.sel1)
d1 = DeviceBuffer(F * n1 * 12, 0); gen1.generate(d1.ptr, 0, F)
d2 = None
if s2 is not None:
    gen2 = device_synth.DeviceTrajectory(s, 0, s.sel2)
    d2 = DeviceBuffer(F * n2 * 12, 0); gen2.generate(d2.ptr, 0, F)
check(lib.pc_device_sync(0))
eng = ContactEngine(s1, s2, 5.0, 2.5, 120, device=0, slots=1)
eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
slot = eng.slots[0]
def T(label, fn):
    t0 = time.perf_counter(); r = fn(); check(lib.pc_device_sync(0)); print("%-28s %9.3f ms" % (label, (time.perf_counter() - t0) * 1e3)); return r
for it in range(3):
    print("--- iteration", it)
    T("scan_device(+acc)", lambda: eng.scan_device(d1.ptr, None if d2 is None else d2.ptr, F, accumulate=True))
    T("fetch_rows", lambda: eng.fetch_rows())
    slot.reserve(F, F * eng.guess[0], F * eng.guess[1], F * eng.guess[2])
    T("pc_scan_device", lambda: check(lib.pc_scan_device(slot.ctx, d1.ptr, None if d2 is None else d2.ptr, F, 3, n1 * 3, n2 * 3, 0, slot.stream)))
    T("pc_accumulate", lambda: check(lib.pc_accumulate(slot.ctx, slot.stream)))
    tot = [ctypes.c_int64() for _ in range(4)]
    T("pc_batch_totals", lambda: lib.pc_batch_totals(slot.ctx, slot.stream, *[ctypes.byref(t) for t in tot]))
    print("   totals", [t.value for t in tot], "guess", eng.guess, "lim", eng.lim)
