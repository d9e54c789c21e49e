"""Wall time of the drop-in object API on the reference's fixture (50 frames, 3097 x 1230 atoms)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from pycontact_b200.core.ContactAnalyzer import Analyzer
from pycontact_b200.io.psf import Topology
z = np.load("tests/golden/rpn11_ubq_input.npz")
topo = Topology([str(s) for s in z["segids"]], z["resids"].astype(np.int64), [str(s) for s in z["resnames"]],
                [str(s) for s in z["names"]], [str(s) for s in z["types"]], np.zeros(len(z["names"])),
                np.ones(len(z["names"])), z["bonds"].astype(np.int64))
coords = z["coords"]
for rep in range(3):
    t0 = time.perf_counter()
    an = Analyzer("a.psf", "a.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ", topology=topo, frames=coords)
    an.runFrameScan(1)
    t1 = time.perf_counter()
    acc = an.runContactAnalysis([0, 0, 1, 1, 0], [0, 0, 1, 1, 0], 1)
    t2 = time.perf_counter()
    n = sum(len(fr) for fr in an.contactResults)
    t3 = time.perf_counter()
    hb = sum(c.hbond_percentage() for c in acc)
    t4 = time.perf_counter()
    print("rep %d: runFrameScan %.1f ms, runContactAnalysis %.1f ms, materialise %d AtomContacts %.1f ms, hbond%% %.1f ms (sum %.1f, %d keys)"
          % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, n, (t3 - t2) * 1e3, (t4 - t3) * 1e3, hb, len(acc)))
