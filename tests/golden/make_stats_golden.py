#!/usr/bin/env python
"""tests/golden/make_stats_golden.py -- golden vectors for the per-contact statistics.

Runs ONLY in the build container: unpickles the reference's golden session
(/root/reference/PyContact/exampleData/defaultsession_py3) with the UNMODIFIED reference classes and calls
their own methods (core/Biochemistry.py:150-213) for two (ns_per_frame, threshold) settings.
Output: tests/golden/rpn11_ubq_stats_golden.npz
"""
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import PyContact.core.Biochemistry  # noqa: E402,F401  (classes named by the pickle)

HERE = os.path.dirname(os.path.abspath(__file__))
d = pickle.load(open("/root/reference/PyContact/exampleData/defaultsession_py3", "rb"))
contacts = d["contacts"]
out = {"score": np.array([c.scoreArray for c in contacts], float),
       "hbonds": np.array([c.hbondFramesScan() for c in contacts], np.int64),
       "mean": np.array([c.mean_score() for c in contacts]),
       "median": np.array([c.median_score() for c in contacts]),
       "hbond_percentage": np.array([c.hbond_percentage() for c in contacts])}
for tag, (ns, thr) in {"a": (1, 0), "b": (0.02, 0.5)}.items():
    out["params_" + tag] = np.array([ns, thr], float)
    out["total_time_" + tag] = np.array([c.total_time(ns, thr) for c in contacts], float)
    out["mean_life_time_" + tag] = np.array([c.mean_life_time(ns, thr) for c in contacts], float)
    out["median_life_time_" + tag] = np.array([c.median_life_time(ns, thr) for c in contacts], float)
    out["life_entries_" + tag] = np.array([len(c.life_time(ns, thr)) for c in contacts], np.int64)
np.savez_compressed(os.path.join(HERE, "rpn11_ubq_stats_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
