"""Kernel time of the small-frame scan per CTA plan (CUDA events on the launching stream, L2 flushed by
input size: 2048 c2 frames are 295 MB)."""
import ctypes, sys
sys.path.insert(0, ".")
import bench
from pycontact_b200 import _lib, device_synth
from pycontact_b200._lib import DeviceBuffer, check
from pycontact_b200.engine import ContactEngine

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
plans = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 3, 4, 1]
w = bench.prepare_workload(name)
s, s1, s2, gm1, gm2 = w["system"], w["s1"], w["s2"], w["gm1"], w["gm2"]
lib = _lib.load()
n1, n2 = s1.n, s2.n
gen1 = device_synth.DeviceTrajectory(s, 0, s.sel1); d1 = DeviceBuffer(F * n1 * 12, 0); gen1.generate(d1.ptr, 0, F)
gen2 = device_synth.DeviceTrajectory(s, 0, s.sel2); d2 = DeviceBuffer(F * n2 * 12, 0); gen2.generate(d2.ptr, 0, F)
check(lib.pc_device_sync(0))
ev = [ctypes.c_void_p(), ctypes.c_void_p()]
for e in ev:
    check(lib.pc_event_create(0, ctypes.byref(e)))
for ctas in plans:
    eng = ContactEngine(s1, s2, 5.0, 2.5, 120, device=0, slots=1)
    eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
    slot = eng.slots[0]
    check(lib.pc_set_small_ctas(slot.ctx, ctas))
    for rep in range(3):
        tot = eng.scan_device(d1.ptr, d2.ptr, F, accumulate=True)
    ms = []
    for rep in range(8):
        check(lib.pc_event_record(ev[0], slot.stream))
        check(lib.pc_scan_device(slot.ctx, d1.ptr, d2.ptr, F, 3, n1 * 3, n2 * 3, 0, slot.stream))
        check(lib.pc_event_record(ev[1], slot.stream))
        check(lib.pc_stream_sync(slot.stream))
        t = ctypes.c_float()
        check(lib.pc_event_elapsed_ms(ev[0], ev[1], ctypes.byref(t)))
        ms.append(t.value)
    st = ctypes.c_uint64(); t4 = [ctypes.c_int64() for _ in range(4)]
    lib.pc_batch_totals(slot.ctx, slot.stream, *[ctypes.byref(t) for t in t4])
    check(lib.pc_last_status(slot.ctx, ctypes.byref(st), None, None, None, None))
    ms.sort()
    print("small_ctas %d: scan %.4f ms median, %.4f min  | totals %s status %d" %
          (ctas, ms[len(ms) // 2], ms[0], [t.value for t in t4], st.value))
    eng.close()
