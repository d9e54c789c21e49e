"""pc_key_stats: kernel time (CUDA events) for a C2-sized score matrix, the host-to-result time of
stats.key_statistics, and the oracle (the reference's per-object Python loops) on a sample of keys."""
import ctypes, sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle import stats_oracle
from pycontact_b200 import _lib, stats
from pycontact_b200._lib import DeviceBuffer, check

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rng = np.random.default_rng(5)
sc = rng.random((K, F)) * np.repeat(rng.random((K, F // 7 + 1)) > 0.45, 7, axis=1)[:, :F]
hb = ((rng.random((K, F)) > 0.8) & (sc > 0)).astype(np.uint32)
lib = _lib.load()
d_sc, d_hb, d_out = DeviceBuffer.from_array(sc), DeviceBuffer.from_array(hb), DeviceBuffer(K * 8 * 8)
ev = [ctypes.c_void_p(), ctypes.c_void_p()]
for e in ev:
    check(lib.pc_event_create(0, ctypes.byref(e)))
ms = []
for rep in range(6):
    check(lib.pc_event_record(ev[0], None))
    check(lib.pc_key_stats(0, d_sc.ptr, d_hb.ptr, K, F, 0.02, 0.5, d_out.ptr, None))
    check(lib.pc_event_record(ev[1], None))
    check(lib.pc_device_sync(0))
    t = ctypes.c_float(); check(lib.pc_event_elapsed_ms(ev[0], ev[1], ctypes.byref(t))); ms.append(t.value)
ms = sorted(ms[1:])
kern = ms[len(ms) // 2]
t0 = time.perf_counter(); st = stats.key_statistics(sc, hb, 0.02, 0.5); api = (time.perf_counter() - t0) * 1e3
ns = 8
t0 = time.perf_counter(); ora = stats_oracle.key_statistics(sc[:ns], hb[:ns], 0.02, 0.5); cpu = (time.perf_counter() - t0) * 1e3 / ns
for k in ("mean", "median", "total_time", "mean_life_time", "median_life_time"):
    np.testing.assert_allclose(st[k][:ns], ora[k], rtol=1e-12)
# the same from the (frame, key) cells (CSR over keys): what an AccumulationTable holds
nz = sc != 0
kp = np.zeros(K + 1, np.int64); np.cumsum(nz.sum(1), out=kp[1:])
kk, ff = np.nonzero(nz)
bufs = [DeviceBuffer.from_array(kp), DeviceBuffer.from_array(ff.astype(np.int32)), DeviceBuffer.from_array(sc[kk, ff]), DeviceBuffer.from_array(hb[kk, ff])]
msr = []
for rep in range(6):
    check(lib.pc_event_record(ev[0], None))
    check(lib.pc_key_stats_rows(0, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, K, F, 0.02, 0.5, d_out.ptr, None))
    check(lib.pc_event_record(ev[1], None))
    check(lib.pc_device_sync(0))
    t = ctypes.c_float(); check(lib.pc_event_elapsed_ms(ev[0], ev[1], ctypes.byref(t))); msr.append(t.value)
kern_rows = sorted(msr[1:])[len(msr[1:]) // 2]
cell_sc, cell_hb, cell_fr = sc[kk, ff], hb[kk, ff], ff.astype(np.int32)
t0 = time.perf_counter(); st2 = stats.key_statistics_rows(kp, cell_fr, cell_sc, cell_hb, F, 0.02, 0.5); api_rows = (time.perf_counter() - t0) * 1e3
assert np.array_equal(st2["median"], st["median"]) and np.array_equal(st2["life_entries"], st["life_entries"])
print("pc_key_stats_rows K=%d F=%d, %d cells (%.0f MB): kernel %.3f ms (%.1f GB/s), through stats.key_statistics_rows %.1f ms"
      % (K, F, kk.size, kk.size * 16 / 1e6, kern_rows, kk.size * 16 / kern_rows / 1e6, api_rows))
bytes_alg = K * F * 12
print("pc_key_stats  K=%d F=%d: kernel %.3f ms (%.1f GB/s of %.0f MB score+hbond matrix), through stats.key_statistics "
      "(H2D + kernel + D2H) %.1f ms; reference loops (oracle, 1 core) %.2f ms per key -> %.0f ms for all keys"
      % (K, F, kern, bytes_alg / kern / 1e6, bytes_alg / 1e6, api, cpu, cpu * K))
