#!/usr/bin/env python
"""C5 sweep (BASELINE.json configs[4]): cutoff x selection size on uniform random clouds (0.1 atoms/A^3,
50 % heavy, N1 = N2), one GPU, frames resident in HBM.  Prints a markdown table:
frames/s, algorithmic GB/s (12 B/atom read + 16 B/contact written), pair evaluations/s, contacts/frame.

    python tools/sweep_c5.py > profiles/r1_c5_sweep.md
"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycontact_b200 import _lib, prepare, synth                      # noqa: E402
from pycontact_b200._lib import DeviceBuffer, check                  # noqa: E402
from pycontact_b200.engine import ContactEngine                      # noqa: E402

SIZES = [1000, 10000, 100000, 1000000]
CUTOFFS = [3.0, 4.0, 5.0, 6.0, 8.0, 10.0, 12.0]


def selection(n, heavy):
    names = ["C" if h else "H" for h in heavy]
    hcsr = (np.zeros(n + 1, np.int64), np.zeros(0, np.int64))
    return prepare.prepare_selection(np.arange(n), names, None, None, None, hcsr)


def main():
    global SIZES, CUTOFFS
    if len(sys.argv) > 2:
        SIZES = [int(float(x)) for x in sys.argv[1].split(",")]
        CUTOFFS = [float(x) for x in sys.argv[2].split(",")]
    lib = _lib.load()
    peak = 6533.5
    print("| N1 = N2 | cutoff A | path | frames/s | alg GB/s | of %.0f | pair evals/s | of FP32 peak | contacts/frame |" % peak)
    print("|---|---|---|---|---|---|---|---|---|")
    for n in SIZES:
        a, b, ha, hb = synth.cloud(n, n, seed=1005 + SIZES.index(n))
        F = int(max(4, min(1024, (96 << 20) // (24 * n))))
        rng = np.random.default_rng(7)
        fa = (a[None] + rng.normal(0, 0.2, (F, n, 3))).astype(np.float32)
        fb = (b[None] + rng.normal(0, 0.2, (F, n, 3))).astype(np.float32)
        d1, d2 = DeviceBuffer.from_array(fa), DeviceBuffer.from_array(fb)
        s1, s2 = selection(n, ha), selection(n, hb)
        for cutoff in CUTOFFS:
            with ContactEngine(s1, s2, cutoff, 2.5, 120, slots=1) as eng:
                tot = eng.scan_device(d1.ptr, d2.ptr, F)             # warm-up, sizes the buffers
                tot = eng.scan_device(d1.ptr, d2.ptr, F)
                reps = 3
                check(lib.pc_device_sync(0))
                t0 = time.perf_counter()
                for _ in range(reps):
                    tot = eng.scan_device(d1.ptr, d2.ptr, F)
                check(lib.pc_device_sync(0))
                dt = (time.perf_counter() - t0) / reps
                path = ctypes.c_int32()
                check(lib.pc_last_status(eng.slots[0].ctx, None, None, None, None, ctypes.byref(path)))
            nc, _, _, ne = tot
            gbs = (24.0 * n * F + 16.0 * nc) / dt / 1e9
            pe = ne / dt
            print("| %d | %.0f | %s | %.3g | %.0f | %.1f %% | %.3g | %.2f %% | %.0f |" % (
                n, cutoff, "small" if path.value == 1 else "large", F / dt, gbs, 100 * gbs / peak, pe,
                100 * pe / (148 * 128 * 1.965e9 / 9), nc / F))
            sys.stdout.flush()
        d1.free(); d2.free()


if __name__ == "__main__":
    main()
