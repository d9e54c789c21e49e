"""baseline/ref_runner.py -- runs the UNMODIFIED reference's CPU path (bench.py --impl reference, cpu_baseline).

``baseline/_ref`` holds ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref`` of the
reference tree (its own setup.py: Cython ``cy_gridsearch`` around ``src/gridsearch.C``); ``install()`` below is
the recipe (``__graft_entry__.build()`` runs it where /root/reference exists, the GPU box uses the installed copy,
which is git-ignored but travels with the snapshot).  Nothing of the reference is copied into this repository.

What is timed is the reference's own code: ``loop_trajectory_grid`` (PyContact/core/multi_trajectory.py:50-222)
with its compiled ``cy_find_contacts``, dispatched like ``run_load_parallel`` does (:423-451): the frame lists
cut by the reference's ``chunks`` (:20-30), one ``apply_async`` per chunk on the reference's ``LoggingPool``
(core/LogPool.py:28-31), results concatenated in order.  ``run_load_parallel`` itself cannot run here -- it opens
the trajectory through MDAnalysis, which is not installed (its import and PyQt5's are satisfied by empty stub
modules; neither is touched by the timed code) -- so the arrays it would have gathered (``sel.positions``,
``sel.indices``, the ``ConvBond`` list, ``name_array`` ...) are built from the synthetic system instead.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "PyContact", "core")) and any(
        f.startswith("cy_gridsearch") and f.endswith(".so")
        for f in os.listdir(os.path.join(REF_DIR, "PyContact", "cy_modules")))


def install(force: bool = False) -> str:
    """pip-install the reference into baseline/_ref from a scratch copy (the build cythonizes into its source tree,
    /root/reference is read-only).  Returns one line describing the outcome."""
    if available() and not force:
        return "baseline/_ref present"
    if not os.path.isdir(SRC):
        return "no %s here; baseline/_ref not built" % SRC
    tmp = "/tmp/pycontact_ref_build"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp)
    shutil.rmtree(REF_DIR, ignore_errors=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", REF_DIR, tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    shutil.rmtree(tmp, ignore_errors=True)
    if res.returncode != 0 or not available():
        return "pip install of the reference failed: " + (res.stderr or res.stdout)[-300:].replace("\n", " ")
    return "installed the reference into baseline/_ref (pip --no-deps: MDAnalysis / PyQt5 / matplotlib are not in this image)"


def _stub_modules():
    """Empty stand-ins for the two third-party imports at the top of core/multi_trajectory.py:6-7 that this image
    lacks.  The timed functions use neither."""
    def mod(name):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    try:
        import MDAnalysis  # noqa: F401
    except Exception:
        mda = mod("MDAnalysis"); ana = mod("MDAnalysis.analysis"); dist = mod("MDAnalysis.analysis.distances")
        mda.analysis = ana; ana.distances = dist
    try:
        import PyQt5.QtCore  # noqa: F401
    except Exception:
        qt = mod("PyQt5"); qtcore = mod("PyQt5.QtCore"); qt.QtCore = qtcore

        class _Signal:
            def __init__(self, *a): pass
            def emit(self, *a): pass
            def connect(self, *a): pass

        class QObject:
            def __init__(self, *a, **k): pass
        qtcore.QObject = QObject
        qtcore.pyqtSignal = lambda *a: _Signal()


_REF = {}


def load():
    """The reference's modules: (multi_trajectory, LogPool, cy_gridsearch)."""
    if _REF:
        return _REF["mt"], _REF["lp"], _REF["cy"]
    if not available():
        raise RuntimeError("baseline/_ref is missing (run baseline.ref_runner.install() where /root/reference exists)")
    _stub_modules()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import PyContact.core.multi_trajectory as mt
    import PyContact.core.LogPool as lp
    import PyContact.cy_modules.cy_gridsearch as cy
    for m in (mt, lp, cy):
        if not os.path.abspath(m.__file__).startswith(REF_DIR):
            raise RuntimeError("%s was not loaded from baseline/_ref" % m.__name__)
    _REF.update(mt=mt, lp=lp, cy=cy)
    return mt, lp, cy


class Bonds:
    """What the reference reads from its ConvBond (core/multi_trajectory.py:33-47, used at :115-155): ``types``
    (tuples of the two atom types per bond) and ``to_indices()`` (index pairs)."""
    __slots__ = ("types", "_idx")

    def __init__(self, types_, idx):
        self.types = types_
        self._idx = idx

    def to_indices(self):
        return self._idx


_NO_BONDS = Bonds([], [])


def bond_table(topo):
    """list over ALL atoms (global index) like ``bonds`` at core/multi_trajectory.py:375-379."""
    ptr, bid = topo.bonds_of_atoms()
    tb = np.asarray(topo.bonds, np.int64)
    types = topo.types
    out = [_NO_BONDS] * topo.n_atoms
    for a in np.nonzero(np.diff(ptr))[0].tolist():
        bs = bid[ptr[a]:ptr[a + 1]]
        out[a] = Bonds([(types[int(tb[b, 0])], types[int(tb[b, 1])]) for b in bs], [tb[b] for b in bs])
    return out


def _suppl(topo, bonds, self_mode):
    names = list(topo.names)
    if self_mode:
        return [bonds, names, list(np.asarray(topo.resids).tolist()), list(topo.segids)]
    return [bonds, names]


def run_chunks(frames1, frames2, idx1, idx2, topo, bonds, self_mode, nproc, cutoff=5.0, hbondcutoff=2.5, hbondcutangle=120):
    """The dispatch of run_load_parallel (core/multi_trajectory.py:423-451) on already gathered frames.
    Returns (list[F] of list[AtomContact], seconds) -- seconds cover chunking, the pool's pickling of arguments and
    results and the workers' loops, like the reference's own wall time; the pool's start-up is outside."""
    mt, lp, _ = load()
    nf = len(frames1)
    i1 = [np.asarray(idx1)] * nf
    i2 = [np.asarray(idx2)] * nf
    suppl = _suppl(topo, bonds, self_mode)
    if nproc <= 1:
        t0 = time.perf_counter()
        out = mt.loop_trajectory_grid(frames1, frames2, i1, i2, [cutoff, hbondcutoff, hbondcutangle], suppl, self_mode)
        return out, time.perf_counter() - t0
    pool = lp.LoggingPool(nproc)
    try:
        t0 = time.perf_counter()
        parts = zip(mt.chunks(frames1, nproc), mt.chunks(frames2, nproc), mt.chunks(i1, nproc), mt.chunks(i2, nproc))
        results = [pool.apply_async(mt.loop_trajectory_grid,
                                    args=(c[0], c[1], c[2], c[3], [cutoff, hbondcutoff, hbondcutangle], suppl, self_mode))
                   for c in parts if len(c[0])]
        out = []
        for r in results:
            out.extend(r.get())
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    return out, dt


def search_only(frames1, frames2, cutoff=5.0):
    """Seconds the reference's compiled cy_find_contacts alone takes on these frames (one process), called with the
    argument shapes of core/multi_trajectory.py:76-79."""
    _, _, cy = load()
    t0 = time.perf_counter()
    npairs = 0
    for a, b in zip(frames1, frames2):
        n1, n2 = len(a), len(b)
        res = cy.cy_find_contacts(np.reshape(a, (1, n1 * 3)), n1, np.reshape(b, (1, n2 * 3)), n2, cutoff)
        npairs += sum(len(r) for r in res)
    return time.perf_counter() - t0, npairs


def frame_summary(frame_contacts):
    """(contacts, hbonds, sum of weights) of one frame's list of AtomContact."""
    return (len(frame_contacts), sum(len(c.hbondinfo) for c in frame_contacts),
            float(sum(c.weight for c in frame_contacts)))
