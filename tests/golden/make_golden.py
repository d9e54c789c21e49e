#!/usr/bin/env python
"""tests/golden/make_golden.py -- generates the committed golden vectors.

Runs ONLY in the build container (needs /root/reference).  It imports the
UNMODIFIED reference package from /root/reference with

  * stub modules for the two uninstalled third-party imports at the top of
    core/multi_trajectory.py:6-7 and core/ContactAnalyzer.py:7-9
    (``MDAnalysis``, ``MDAnalysis.analysis.distances``, ``PyQt5.QtCore``), and
  * ``PyContact.cy_modules.cy_gridsearch`` provided by the reference's own
    gridsearch.C compiled into oracle/_ref (oracle/Makefile) behind a wrapper with the
    signature/return shape of cy_gridsearch.pyx:31-35,

and then calls the reference's ``loop_trajectory_grid`` (core/multi_trajectory.py:50-222)
and ``Analyzer.analyze_contactResultsWithMaps`` (core/ContactAnalyzer.py:502-597) on

  A. the rpn11_ubq fixture, ``segid RN11`` vs ``segid UBQ`` (tests/test_basic.py:32-43)
  B. the same fixture, ``segid RN11`` vs ``self``          (tests/test_basic.py:45-51)
  C. a small synthetic system from pycontact_b200.synth (seeded), normal + self mode

It also converts the reference's pickled session exampleData/defaultsession_py3 (written
by core/DataHandler.py:20-28) into arrays.  Outputs (all under tests/golden/):

  rpn11_ubq_input.npz      topology arrays + float32 coordinates of all 50 frames
  rpn11_ubq_golden.npz     case A: per-contact / per-H-bond / accumulated arrays, from BOTH the
                           live reference run and the pickled session (they must agree)
  rpn11_self_golden.npz    case B: full arrays for 3 frames + per-frame counts/checksums for all
  synth_small_golden.npz   case C
"""
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import native  # noqa: E402
from pycontact_b200.io.psf import read_psf  # noqa: E402
from pycontact_b200.io.dcd import DCDReader  # noqa: E402
from pycontact_b200.io.selection import select_atoms  # noqa: E402


def install_stubs():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    mda = mod("MDAnalysis")
    ana = mod("MDAnalysis.analysis")
    dist = mod("MDAnalysis.analysis.distances")
    mda.analysis = ana
    ana.distances = dist
    qt = mod("PyQt5")
    qtcore = mod("PyQt5.QtCore")
    qt.QtCore = qtcore

    class _Signal:
        def __init__(self, *a):
            pass

        def emit(self, *a):
            pass

        def connect(self, *a):
            pass

    class QObject:
        def __init__(self, *a, **k):
            pass

    qtcore.QObject = QObject
    qtcore.pyqtSignal = lambda *a: _Signal()

    # the compiled reference search behind the pyx's Python signature
    cy = mod("PyContact.cy_modules.cy_gridsearch")

    def cy_find_contacts(xyz1, natoms1, xyz2, natoms2, r):
        row_ptr, cols = native.ref_find_contacts(np.asarray(xyz1[0], np.float32),
                                                 np.asarray(xyz2[0], np.float32), r)
        return native.as_list_of_lists(row_ptr, cols, natoms2)
    cy.cy_find_contacts = cy_find_contacts


class RefBond:
    """Object with the attributes the reference reads from ConvBond
    (core/multi_trajectory.py:33-47): .types and .to_indices()."""

    def __init__(self, types_, indices):
        self.types = types_
        self.indices = indices

    def to_indices(self):
        return self.indices


def ref_bonds(topo):
    ptr, bid = topo.bonds_of_atoms()
    out = []
    for a in range(topo.n_atoms):
        bs = bid[ptr[a]:ptr[a + 1]]
        out.append(RefBond([(topo.types[int(topo.bonds[b, 0])], topo.types[int(topo.bonds[b, 1])]) for b in bs],
                           [np.array([topo.bonds[b, 0], topo.bonds[b, 1]]) for b in bs]))
    return out


def contacts_to_arrays(frames):
    """list[F] of list[AtomContact]  ->  flat arrays + per-frame pointers."""
    fptr = [0]
    idx1, idx2, dist, wt, hptr = [], [], [], [], [0]
    hd, ha, hh, hdist, hang = [], [], [], [], []
    for fr in frames:
        for c in fr:
            idx1.append(c.idx1); idx2.append(c.idx2); dist.append(c.distance); wt.append(c.weight)
            for hb in c.hbondinfo:
                hd.append(hb.donorIndex); ha.append(hb.acceptorIndex); hh.append(hb.hydrogenIndex)
                hdist.append(hb.acceptorHydrogenDistance); hang.append(hb.angle)
            hptr.append(len(hd))
        fptr.append(len(idx1))
    return dict(frame_ptr=np.asarray(fptr, np.int64), idx1=np.asarray(idx1, np.int32),
                idx2=np.asarray(idx2, np.int32), distance=np.asarray(dist, np.float64),
                weight=np.asarray(wt, np.float64), hb_ptr=np.asarray(hptr, np.int64),
                hb_donor=np.asarray(hd, np.int32), hb_acceptor=np.asarray(ha, np.int32),
                hb_hydrogen=np.asarray(hh, np.int32), hb_distance=np.asarray(hdist, np.float64),
                hb_angle=np.asarray(hang, np.float64))


def accumulated_to_arrays(acc_list, prefix="acc_"):
    K = len(acc_list)
    F = len(acc_list[0].scoreArray) if K else 0
    score = np.zeros((K, F))
    hb = np.zeros((K, F), np.int64)
    ncontrib = np.zeros((K, F), np.int64)
    for k, a in enumerate(acc_list):
        score[k] = np.asarray(a.scoreArray, float)
        hb[k] = a.hbondFramesScan()
        ncontrib[k] = [len(x) for x in a.contributingAtoms]
    return {
        prefix + "key1": np.asarray([[str(x) for x in a.key1] for a in acc_list]),
        prefix + "key2": np.asarray([[str(x) for x in a.key2] for a in acc_list]),
        prefix + "title": np.asarray([a.title for a in acc_list]),
        prefix + "score": score, prefix + "hbonds": hb, prefix + "ncontrib": ncontrib,
        prefix + "bb1": np.asarray([a.bb1 for a in acc_list], float),
        prefix + "sc1": np.asarray([a.sc1 for a in acc_list], float),
        prefix + "bb2": np.asarray([a.bb2 for a in acc_list], float),
        prefix + "sc2": np.asarray([a.sc2 for a in acc_list], float),
        prefix + "hbond_percentage": np.asarray([a.hbond_percentage() for a in acc_list], float),
    }


def run_reference(topo, coords, sel1, sel2, cutoff, hbc, hba, self_mode, maps, backbone):
    from PyContact.core.multi_trajectory import loop_trajectory_grid
    from PyContact.core.ContactAnalyzer import Analyzer
    bonds = ref_bonds(topo)
    names = list(topo.names)
    resid = [int(r) for r in topo.resids]
    segids = list(topo.segids)
    F = coords.shape[0]
    s1 = [np.ascontiguousarray(coords[f, sel1]) for f in range(F)]
    s2 = [np.ascontiguousarray(coords[f, sel2]) for f in range(F)]
    i1 = [np.asarray(sel1)] * F
    i2 = [np.asarray(sel2)] * F
    suppl = [bonds, names, resid, segids] if self_mode else [bonds, names]
    res = loop_trajectory_grid(s1, s2, i1, i2, [cutoff, hbc, hba], suppl, self_mode)
    an = Analyzer("psf", "dcd", cutoff, hbc, hba, "sel1", "self" if self_mode else "sel2")
    an.setTrajectoryData(list(topo.resnames), resid, names, segids, [int(b) for b in backbone], "sel1", "sel2")
    an.contactResults = res
    acc = an.runContactAnalysis(maps[0], maps[1], 1)
    return res, acc


def checksum(c, f):
    lo, hi = c["frame_ptr"][f], c["frame_ptr"][f + 1]
    a = c["idx1"][lo:hi].astype(np.int64)
    b = c["idx2"][lo:hi].astype(np.int64)
    return int(((a * 1000003 + b) % 2147483647).sum())


def main():
    assert os.path.isdir(REF), "needs the reference tree"
    native.build()
    install_stubs()
    sys.path.insert(0, REF)

    psf = os.path.join(REF, "PyContact/exampleData/rpn11_ubq.psf")
    dcd = os.path.join(REF, "PyContact/exampleData/rpn11_ubq.dcd")
    topo = read_psf(psf)
    traj = DCDReader(dcd)
    F = len(traj)
    coords = np.stack([traj.frame(f) for f in range(F)])
    backbone = select_atoms(topo, "backbone")

    # --- pickled session of the reference (known-answer data) ---------------------------
    with open(os.path.join(REF, "PyContact/exampleData/defaultsession_py3"), "rb") as fh:
        sess = pickle.load(fh)
    t_resname, t_resid, t_name, t_segids, t_backbone, t_s1, t_s2 = sess["trajectory"]
    assert list(t_name) == list(topo.names), "PSF reader disagrees with the session's name array"
    assert [int(r) for r in t_resid] == [int(r) for r in topo.resids]
    assert list(t_resname) == list(topo.resnames)
    assert list(t_segids) == list(topo.segids)
    assert [int(b) for b in t_backbone] == [int(b) for b in backbone], "backbone selection differs"
    pk_contacts = contacts_to_arrays(sess["analyzer"][-1])
    pk_acc = accumulated_to_arrays(sess["contacts"], "acc_")
    maps = sess["maps"]
    print("session:", sess["analyzer"][:7], "maps", maps, "contacts", len(pk_contacts["idx1"]),
          "hbonds", len(pk_contacts["hb_donor"]), "keys", len(sess["contacts"]))

    np.savez_compressed(
        os.path.join(HERE, "rpn11_ubq_input.npz"),
        coords=coords, names=np.asarray(topo.names), types=np.asarray(topo.types),
        resnames=np.asarray(topo.resnames), segids=np.asarray(topo.segids),
        resids=topo.resids.astype(np.int32), bonds=topo.bonds.astype(np.int32),
        backbone=backbone.astype(np.int32))

    # --- case A: live reference run must reproduce the pickle ---------------------------
    sel1 = select_atoms(topo, "segid RN11")
    sel2 = select_atoms(topo, "segid UBQ")
    res, acc = run_reference(topo, coords, sel1, sel2, 5.0, 2.5, 120, False, [[0, 0, 1, 1, 0]] * 2, backbone)
    live = contacts_to_arrays(res)
    for k in ("frame_ptr", "idx1", "idx2", "hb_ptr", "hb_donor", "hb_acceptor", "hb_hydrogen"):
        assert np.array_equal(live[k], pk_contacts[k]), "live reference != pickled session in " + k
    for k in ("distance", "weight", "hb_distance", "hb_angle"):
        assert np.allclose(live[k], pk_contacts[k], rtol=1e-12, atol=0), k
    live_acc = accumulated_to_arrays(acc, "acc_")
    assert np.array_equal(live_acc["acc_key1"], pk_acc["acc_key1"])
    assert np.allclose(live_acc["acc_score"], pk_acc["acc_score"], rtol=1e-12)
    assert float(live_acc["acc_hbond_percentage"].sum()) == 676.0
    raw = []
    for f in range(F):
        rp, cols = native.ref_find_contacts(coords[f, sel1], coords[f, sel2], 5.0)
        raw.append(int(rp[-1]))
    out = dict(pk_contacts)
    out.update(pk_acc)
    out.update(sel1=sel1.astype(np.int32), sel2=sel2.astype(np.int32), n_raw=np.asarray(raw, np.int64),
               config=np.asarray([5.0, 2.5, 120.0]), map1=np.asarray(maps[0], np.int32),
               map2=np.asarray(maps[1], np.int32))
    np.savez_compressed(os.path.join(HERE, "rpn11_ubq_golden.npz"), **out)
    print("case A ok: frames", F, "contacts", len(live["idx1"]), "raw pairs", sum(raw),
          "keys", len(acc), "sum hb%", live_acc["acc_hbond_percentage"].sum())

    # --- case B: self mode -------------------------------------------------------------------
    res, acc = run_reference(topo, coords, sel1, sel1, 5.0, 2.5, 120, True, [[0, 0, 1, 1, 0]] * 2, backbone)
    live = contacts_to_arrays(res)
    acc_arr = accumulated_to_arrays(acc, "acc_")
    keep = [0, F // 2, F - 1]
    sub = contacts_to_arrays([res[f] for f in keep])
    out = {"full_" + k: v for k, v in sub.items()}
    out.update(full_frames=np.asarray(keep, np.int32),
               frame_counts=np.diff(live["frame_ptr"]),
               frame_hbonds=np.asarray([sum(len(c.hbondinfo) for c in fr) for fr in res], np.int64),
               frame_weight_sum=np.asarray([sum(c.weight for c in fr) for fr in res], float),
               frame_checksum=np.asarray([checksum(live, f) for f in range(F)], np.int64),
               sel1=sel1.astype(np.int32), config=np.asarray([5.0, 2.5, 120.0]))
    # accumulated: keep everything except the (K,F) ncontrib which the score mask implies
    out.update(acc_arr)
    np.savez_compressed(os.path.join(HERE, "rpn11_self_golden.npz"), **out)
    print("case B ok: contacts", len(live["idx1"]), "hbonds", len(live["hb_donor"]),
          "sum w", live["weight"].sum(), "keys", len(acc), "sum hb%", acc_arr["acc_hbond_percentage"].sum())

    # --- case C: small synthetic system ---------------------------------------------------------
    from pycontact_b200 import synth
    sysm = synth.make_system("tiny", seed=7)
    coords_s = synth.frames_host(sysm, range(6))
    out = {}
    for tag, self_mode in (("ab", False), ("self", True)):
        s1 = sysm.sel1
        s2 = sysm.sel1 if self_mode else sysm.sel2
        res, acc = run_reference(sysm.topology, coords_s, s1, s2, 5.0, 2.5, 120, self_mode,
                                 [[0, 0, 1, 1, 1]] * 2, sysm.backbone)
        arr = contacts_to_arrays(res)
        out.update({tag + "_" + k: v for k, v in arr.items()})
        out.update({tag + "_" + k: v for k, v in accumulated_to_arrays(acc, "acc_").items()})
        print("case C", tag, "contacts", len(arr["idx1"]), "hbonds", len(arr["hb_donor"]), "keys", len(acc))
    out["coords"] = coords_s
    np.savez_compressed(os.path.join(HERE, "synth_small_golden.npz"), **out)


if __name__ == "__main__":
    main()
