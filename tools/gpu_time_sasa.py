"""pc_sasa_device / pc_within_device: CUDA-event time for frames resident in HBM, the host-to-result time of
sasa_frames, and the compiled reference (oracle/_ref, 1 core, one frame at a time like cy_sasa) on a sample.

    python tools/gpu_time_sasa.py [natoms] [frames]        # default: a 1230-atom selection (UBQ-sized), 2000 frames
"""
import ctypes, sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle import sasa_oracle as so
from pycontact_b200 import _lib
from pycontact_b200._lib import DeviceBuffer, check
from pycontact_b200.cy_modules.cy_gridsearch import sasa_frames, find_within_frames

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1230
F = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rng = np.random.default_rng(9)
L = (n / 0.1) ** (1 / 3)
base = (rng.random((n, 3)) * L).astype(np.float32)
coords = (base[None] + rng.normal(0, 0.3, (F, n, 3))).astype(np.float32)
rad = rng.choice(np.array([1.0, 1.5, 1.4, 1.3, 1.9], np.float32), n).astype(np.float32)
rl = np.array([0], np.int32)
PAIRDIST = 6.8
lib = _lib.load()
d_xyz = DeviceBuffer.from_array(coords)
d_tot = DeviceBuffer(F * 4)
d_ev = DeviceBuffer(8)
ev = [ctypes.c_void_p(), ctypes.c_void_p()]
for e in ev:
    check(lib.pc_event_create(0, ctypes.byref(e)))


def timed(fn, reps=5):
    ms = []
    for _ in range(reps):
        check(lib.pc_event_record(ev[0], None))
        fn()
        check(lib.pc_event_record(ev[1], None))
        check(lib.pc_device_sync(0))
        t = ctypes.c_float(); check(lib.pc_event_elapsed_ms(ev[0], ev[1], ctypes.byref(t))); ms.append(t.value)
    return sorted(ms[1:])[len(ms[1:]) // 2]


check(lib.pc_memset(0, d_ev.ptr, 0, 8, None))
k0 = lib.pc_launch_count()
ms = timed(lambda: check(lib.pc_sasa_device(0, d_xyz.ptr, F, n, 3 * n, _lib.ptr(rad), PAIRDIST, 50, 1.4, 1, 0, _lib.ptr(rl),
                                            d_tot.ptr, None, d_ev.ptr, None)))
launches = (lib.pc_launch_count() - k0) // 5
evals = int(d_ev.download(np.uint64, 1)[0]) // 5
tot = d_tot.download(np.float32, F)
t_apis = []
for _ in range(3):
    t0 = time.perf_counter(); api = sasa_frames(coords, rad, PAIRDIST, 50, 1.4, 1, 0, rl); t_apis.append((time.perf_counter() - t0) * 1e3)
t_api = min(t_apis)
print('sasa_frames calls: ' + ', '.join('%.1f ms' % t for t in t_apis))
assert np.array_equal(api, tot)
ns = max(1, min(F, int(2e5 // max(n, 1)) + 1))
t0 = time.perf_counter()
ref = np.array([so.ref_sasa(coords[f], rad, PAIRDIST, 50, 1.4, 1, 0, rl) for f in range(ns)], np.float32)
t_ref = (time.perf_counter() - t0) * 1e3 / ns
assert np.array_equal(ref, tot[:ns]), "GPU totals differ from the reference"
print("pc_sasa_device natoms=%d frames=%d: %.3f ms (%d launches) = %.0f frames/s, %.3g atoms/s, %.3g point-sphere tests/s; "
      "input %.1f GB/s; through sasa_frames (H2D + kernels + D2H) %.1f ms; reference sasa_grid (1 core) %.2f ms per frame "
      "-> %.0fx per frame on the device, %.0fx through the host API"
      % (n, F, ms, launches, F / ms * 1e3, F * n / ms * 1e3, evals / ms * 1e3, F * n * 12 / ms / 1e6, t_api, t_ref,
         t_ref * F / ms, t_ref * F / t_api))

flg = (rng.random(n) < 0.5).astype(np.int32)
oth = (1 - flg).astype(np.int32)
d_f, d_o, d_res = DeviceBuffer.from_array(flg), DeviceBuffer.from_array(oth), DeviceBuffer(F * n * 4)
ms = timed(lambda: check(lib.pc_within_device(0, d_xyz.ptr, F, n, 3 * n, d_f.ptr, d_o.ptr, 5.0, d_res.ptr, None)))
t0 = time.perf_counter()
for f in range(ns):
    so.ref_find_within(coords[f], flg, oth, 5.0)
t_ref = (time.perf_counter() - t0) * 1e3 / ns
got = d_res.download(np.int32, F * n).reshape(F, n)
assert np.array_equal(got[0], so.ref_find_within(coords[0], flg, oth, 5.0))
print("pc_within_device natoms=%d frames=%d r=5: %.3f ms = %.0f frames/s (%.1f GB/s in+out); reference find_within %.3f ms per frame -> %.0fx"
      % (n, F, ms, F / ms * 1e3, F * n * 16 / ms / 1e6, t_ref, t_ref * F / ms))
