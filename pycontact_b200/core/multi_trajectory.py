"""Drop-ins for ``PyContact.core.multi_trajectory``: ``loop_trajectory_grid`` (:50-222),
``run_load_parallel`` (:354-455), ``chunks`` (:20-30), ``weight_function`` (:15-17) and
``ConvBond`` (:33-47), with the frame loop on the GPU.

``loop_trajectory_grid`` keeps the reference's argument list -- lists of per-frame
selection coordinates and index arrays, ``config = [cutoff, hbondcutoff, hbondcutangle]``,
``suppl = [bonds, name_array(, resid_array, segids)]`` -- and returns one list of
``AtomContact`` per frame.  Two quirks of the reference are kept on purpose (SURVEY.md
App. A.6 Q1): every returned ``AtomContact.frame`` is 0, and the index arrays of the chunk's
FIRST frame are used for all frames (the reference's ``frame`` counter never advances).
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from .. import prepare
from ..engine import ContactEngine
from ..results import LazyContactResults
from ..scheduler import ArraySource, chunks, run_scan   # noqa: F401  (chunks is part of this module's API)
from .Biochemistry import *                              # noqa: F401,F403  (the reference star-imports the types)


def weight_function(value):
    """Sigmoid contact score; the CUDA path evaluates the same expression in float64."""
    return 1.0 / (1.0 + np.exp(5.0 * (value - 4.0)))


class ConvBond(object):
    """Plain-data copy of an MDAnalysis bond group: ``types`` (atom-type tuples) and ``indices``."""

    def __init__(self, bonds):
        self.types = [deepcopy(b.type) for b in bonds]
        self.indices = [deepcopy(b.indices) for b in bonds]

    def to_indices(self):
        return self.indices


def loop_trajectory_grid(sel1c, sel2c, indices1, indices2, config, suppl, selfInteraction, *, device=0,
                         batch_frames=None):
    cutoff, hbondcutoff, hbondcutangle = config
    bonds, name_array = suppl[0], suppl[1]
    resid_array, segids = (suppl[2], suppl[3]) if selfInteraction else ([], [])
    nframes = min(len(sel1c), len(sel2c))
    if nframes == 0:
        return []
    idx1 = np.asarray(indices1[0], np.int64)          # Q1: indices of the chunk's first frame
    idx2 = np.asarray(indices2[0], np.int64)
    if idx1.size == 0 or idx2.size == 0:
        return [[] for _ in range(nframes)]
    hcsr = prepare.hydrogen_csr_from_bond_objects(bonds, name_array)
    seg_codes = prepare.string_codes(segids) if selfInteraction else None
    s1 = prepare.prepare_selection(idx1, name_array, resid_array if selfInteraction else None, seg_codes, None, hcsr)
    same = selfInteraction and idx1.size == idx2.size and np.array_equal(idx1, idx2)
    s2 = s1 if same else prepare.prepare_selection(idx2, name_array, resid_array if selfInteraction else None,
                                                   seg_codes, None, hcsr)
    # "self" in the reference means sel2 := sel1 (two separate coordinate copies, same atoms): the
    # device's self mode reads one buffer; anything else is a two-selection scan with the filter on
    if selfInteraction and not same:
        raise NotImplementedError("selfInteraction with two different index sets")
    eng = ContactEngine(s1, None if same else s2, cutoff, hbondcutoff, hbondcutangle, device=device)
    try:
        src = ArraySource(sel1c[:nframes], None if same else sel2c[:nframes])
        traj, _, _ = run_scan([eng], src, batch_frames=batch_frames, order_keys=True, keep_contacts=True)
    finally:
        eng.close()
    return LazyContactResults(traj, frame_label=0)


def run_load_parallel(nproc, psf, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text, **kw):
    """Load, scan on up to ``nproc`` GPUs, return
    ``[allContacts, resname_array, resid_array, name_array, segids, backbone]`` (:455)."""
    from .ContactAnalyzer import Analyzer
    an = Analyzer(psf, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text, **kw)
    res = an.analyze_psf_dcd_grid(psf, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text, nproc=nproc)
    res = LazyContactResults(res.traj, frame_label=0)          # Q1: the pool path reports frame 0
    return [res, an.resname_array, an.resid_array, an.name_array, an.segids, an.backbone]
