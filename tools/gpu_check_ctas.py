import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import bench
from pycontact_b200 import _lib, device_synth
from pycontact_b200._lib import DeviceBuffer, check
from pycontact_b200.engine import ContactEngine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
w = bench.prepare_workload(name)
s, s1, s2, gm1, gm2 = w["system"], w["s1"], w["s2"], w["gm1"], w["gm2"]
lib = _lib.load()
n1, n2 = s1.n, s2.n
gen1 = device_synth.DeviceTrajectory(s, 0, s.sel1); d1 = DeviceBuffer(F * n1 * 12, 0); gen1.generate(d1.ptr, 0, F)
gen2 = device_synth.DeviceTrajectory(s, 0, s.sel2); d2 = DeviceBuffer(F * n2 * 12, 0); gen2.generate(d2.ptr, 0, F)
check(lib.pc_device_sync(0))
for ctas in (1, 2, 0):
    for fused in (1, 0):
        eng = ContactEngine(s1, s2, 5.0, 2.5, 120, device=0, slots=1)
        eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
        lib.pc_set_small_ctas(eng.slots[0].ctx, ctas)
        lib.pc_set_fused_accumulate(eng.slots[0].ctx, fused)
        for rep in range(2):
            tot = eng.scan_device(d1.ptr, d2.ptr, F, accumulate=True)
        rows, rs, rc = eng.fetch_rows(0)
        st = ctypes.c_uint64(); check(lib.pc_last_status(eng.slots[0].ctx, ctypes.byref(st), None, None, None, None))
        print("ctas", ctas, "fused", fused, "totals", tot, "status", st.value, "rows sum score %.6f" % rows["score"].sum(), "lim", eng.lim, "small_ctas", eng.small_ctas)
        eng.close()
