"""``Analyzer`` -- the drop-in for ``PyContact.core.ContactAnalyzer.Analyzer``
(PyContact/core/ContactAnalyzer.py:19-613) with the per-frame work on the GPU.

Same constructor, same methods, same attributes; what differs is where the work runs:

* ``runFrameScan(nproc)``          (:48-62)  frames are streamed through CUDA in batches;
  ``nproc`` = how many GPUs of this host to shard the frames over (the reference's
  ``nproc`` worker processes; capped at the GPUs present).
* ``runContactAnalysis(map1, map2)`` (:64-70, :502-597)  per-frame reduce-by-key on the
  device, then ``AccumulatedContact`` objects with lazy per-frame arrays.
* ``contactResults`` is an array-backed sequence that builds ``AtomContact`` objects on
  access (``pycontact_b200.results.LazyContactResults``).

Topology / trajectory loading -- MDAnalysis in the reference (:282-313) -- is done by the
small PSF / DCD readers in ``pycontact_b200.io``; pass ``topology=`` and ``frames=`` to run
on arrays instead of files (synthetic systems, other loaders).
"""
from __future__ import annotations

import time
from copy import deepcopy
from typing import List, Optional

import numpy as np

from .. import prepare
from ..engine import ContactEngine
from ..results import AccumulationTable, LazyContactResults, TrajectoryContacts
from ..scheduler import ArraySource, DCDSource, run_scan
from .Biochemistry import (AccumulatedContact, AccumulationMapIndex, AtomContact, HydrogenBond,
                           HydrogenBondAtoms, TempContactAccumulate)


class _Signal:
    """Stand-in for the Qt signal ``frameUpdate = pyqtSignal(float)`` (:21): connect callables,
    emit progress in [0, 1].  PyQt5 is a GUI dependency and is not imported by the hot path."""

    def __init__(self):
        self._slots = []

    def connect(self, fn):
        self._slots.append(fn)

    def disconnect(self, fn=None):
        self._slots = [s for s in self._slots if fn is not None and s is not fn]

    def emit(self, value):
        for s in list(self._slots):
            s(value)


def _device_count() -> int:
    from .. import _lib
    return int(_lib.load().pc_device_count())


class Analyzer(object):
    """Performs a contact search and analyzes the results (GPU)."""

    def __init__(self, psf, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text, *,
                 topology=None, frames=None, devices=None, batch_frames=None, keep_contacts=True):
        self.frameUpdate = _Signal()
        self.psf = psf
        self.dcd = dcd
        self.cutoff = cutoff
        self.hbondcutoff = hbondcutoff
        self.hbondcutangle = hbondcutangle
        self.sel1text = sel1text
        self.sel2text = sel2text
        self.lastMap1 = []
        self.lastMap2 = []
        self.contactResults = []
        self.resname_array = []
        self.resid_array = []
        self.name_array = []
        self.segids = []
        self.backbone = []
        self.finalAccumulatedContacts = []
        self.bonds = []
        self.totalFrameNumber = 0
        self.currentFrameNumber = 0
        self.analysis_state = False
        self.totalFramesToProcess = 1
        # GPU-path state
        self._topology = topology
        self._frames = frames
        self._devices = devices
        self._batch_frames = batch_frames
        self._keep_contacts = keep_contacts
        self._sel = None            # (SelectionTopology, SelectionTopology)
        self._traj: Optional[TrajectoryContacts] = None
        self._table: Optional[AccumulationTable] = None
        self.scanStats = {}

    # ---- reference API --------------------------------------------------------------------------
    def runFrameScan(self, nproc):
        """Performs a contact search on ``nproc`` GPUs."""
        try:
            self.contactResults = self.analyze_psf_dcd_grid(self.psf, self.dcd, self.cutoff, self.hbondcutoff,
                                                            self.hbondcutangle, self.sel1text, self.sel2text,
                                                            nproc=nproc)
        except Exception:
            raise Exception          # the reference turns every failure into a bare Exception (:61-62)

    def runContactAnalysis(self, map1, map2, nproc=1):
        """Accumulates the scanned contacts under the two maps; ``nproc`` is ignored, like in the
        reference (:64-70, App. A.6 Q8)."""
        self.finalAccumulatedContacts = self.analyze_contactResultsWithMaps(self.contactResults, map1, map2)
        self.lastMap1 = map1
        self.lastMap2 = map2
        return deepcopy(self.finalAccumulatedContacts)

    def setTrajectoryData(self, resname_array, resid_array, name_array, segids, backbone, sel1text, sel2text):
        self.resname_array = resname_array
        self.resid_array = resid_array
        self.name_array = name_array
        self.segids = segids
        self.backbone = backbone
        self.sel1text = sel1text
        self.sel2text = sel2text

    def getTrajectoryData(self):
        return [self.resname_array, self.resid_array, self.name_array, self.segids, self.backbone,
                self.sel1text, self.sel2text]

    def getFilePaths(self):
        return self.psf, self.dcd

    @staticmethod
    def find_between(s, first, last):
        a = s.find(first)
        if a < 0:
            return ""
        a += len(first)
        b = s.find(last, a)
        return "" if b < 0 else s[a:b]

    @staticmethod
    def weight_function(value):
        """Sigmoid contact score (:102-105); the GPU path evaluates the same expression in float64."""
        return 1.0 / (1.0 + np.exp(5.0 * (value - 4.0)))

    def _atom_props(self, idx):
        return (idx, self.name_array[idx], self.resid_array[idx], self.resname_array[idx], self.segids[idx])

    def makeKeyArraysFromMaps(self, map1, map2, contact):
        """Key arrays of one contact: the selected properties of atom 1 / atom 2, "none" elsewhere."""
        p1, p2 = self._atom_props(contact.idx1), self._atom_props(contact.idx2)
        return [[p1[k] if map1[k] == 1 else "none" for k in range(5)],
                [p2[k] if map2[k] == 1 else "none" for k in range(5)]]

    @staticmethod
    def makeKeyFromKeyArrays(key1, key2):
        tok = AccumulationMapIndex.mapping
        half = lambda key: "".join(tok[k] + str(v) for k, v in enumerate(key) if v != "none")
        return half(key1) + "-" + half(key2)

    def makeKeyArraysFromKey(self, key):
        """Inverse of makeKeyFromKeyArrays; every component comes back as a string.  Tokens are
        located left to right in map order (a property value containing a later token is cut
        there, like in the reference, App. A.6 Q7)."""
        out = []
        tok = AccumulationMapIndex.mapping
        for half in key.split("-")[:2]:
            arr = []
            for k, t in enumerate(tok):
                a = half.find(t)
                if a < 0:
                    arr.append("none")
                    continue
                a += len(t)
                rest = half[a:]
                cut = len(rest)
                for t2 in tok[k + 1:]:
                    b = rest.find(t2)
                    if b >= 0:
                        cut = b
                        break
                arr.append(rest[:cut] if rest[:cut] != "" else "none")
            out.append(arr)
        return out

    @staticmethod
    def make_single_title(key):
        get = lambda k: "" if key[k] == "none" else str(key[k])
        parts = ["%s%s" % (get(AccumulationMapIndex.resname), get(AccumulationMapIndex.resid))]
        for k in (AccumulationMapIndex.index, AccumulationMapIndex.name, AccumulationMapIndex.segid):
            if get(k):
                parts.append("%s %s" % (AccumulationMapIndex.mapping[k], get(k)))
        return ", ".join(p for p in parts if p)

    # ---- frame scan --------------------------------------------------------------------------------
    def _load_topology(self, psf):
        if self._topology is not None:
            return self._topology
        from ..io.psf import read_psf
        if not str(psf).lower().endswith(".psf"):
            raise NotImplementedError("only PSF topologies are read natively; pass topology= for other formats")
        return read_psf(psf)

    def _frame_source(self, dcd, idx1, idx2):
        if self._frames is not None:
            fr = self._frames
            if isinstance(fr, (tuple, list)) and len(fr) == 2 and not isinstance(fr[0], (int, float)) \
                    and np.asarray(fr[0][0]).ndim >= 2 and len(fr[0]) == len(fr[1]) and \
                    np.asarray(fr[0][0]).shape[0] == len(idx1):
                # (frames_sel1, frames_sel2): already gathered per selection
                return ArraySource(fr[0], None if idx2 is None else fr[1])
            arr = np.asarray(fr)
            if arr.ndim != 3:
                raise ValueError("frames= must be (F, natoms, 3) or a pair of per-selection frame lists")
            a = np.ascontiguousarray(arr[:, idx1], np.float32)
            b = None if idx2 is None else np.ascontiguousarray(arr[:, idx2], np.float32)
            return ArraySource(a, b)
        from ..io.xtc import open_trajectory
        return DCDSource(open_trajectory(dcd), idx1, idx2)      # (.dcd or .xtc: the readers share one interface)

    def analyze_psf_dcd_grid(self, psf, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text, nproc=1):
        """Load topology + trajectory, scan all frames on the GPU(s) (:278-500)."""
        from ..io.selection import select_atoms
        topo = self._load_topology(psf)
        self.resname_array = list(topo.resnames)
        self.resid_array = [int(r) for r in topo.resids]
        self.name_array = list(topo.names)
        self.segids = list(topo.segids)
        self.backbone = [int(i) for i in select_atoms(topo, "backbone")]
        if "around" in sel1text or "around" in sel2text:
            return self._scan_dynamic_selections(topo, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text)
        self_mode = sel2text == "self"
        idx1 = select_atoms(topo, sel1text)
        idx2 = idx1 if self_mode else select_atoms(topo, sel2text)
        if len(idx1) == 0 or len(idx2) == 0:
            raise Exception("empty atom selection")                      # :312-313
        hcsr = prepare.hydrogen_csr_from_bond_table(topo.names, topo.types, topo.bonds)
        bb = np.zeros(topo.n_atoms, bool)
        bb[self.backbone] = True
        seg_codes = prepare.string_codes(topo.segids)
        s1 = prepare.prepare_selection(idx1, topo.names, topo.resids, seg_codes, bb, hcsr)
        s2 = s1 if self_mode else prepare.prepare_selection(idx2, topo.names, topo.resids, seg_codes, bb, hcsr)
        self._sel = (s1, s2)
        source = self._frame_source(dcd, idx1, None if self_mode else idx2)
        self.totalFrameNumber = source.n_frames
        devices = self._devices
        if devices is None:
            devices = list(range(max(1, min(int(nproc), _device_count()))))
        start = time.time()
        engines = [ContactEngine(s1, None if self_mode else s2, cutoff, hbondcutoff, hbondcutangle, device=d)
                   for d in devices]
        try:
            traj, _, stats = run_scan(engines, source, batch_frames=self._batch_frames, order_keys=True,
                                      accumulate=False, keep_contacts=True)
        finally:
            for e in engines:
                e.close()
        stats["seconds"] = time.time() - start
        self.scanStats = stats
        self.currentFrameNumber = max(0, source.n_frames - 1)
        self._traj = traj
        return LazyContactResults(traj)

    def _all_frames(self, dcd, natoms):
        """(F, natoms, 3) float32 of the whole trajectory (the reference's serial loop walks u.trajectory)."""
        if self._frames is not None:
            arr = np.asarray(self._frames)
            if arr.ndim != 3 or arr.shape[1] != natoms:
                raise ValueError("selections with 'around' need frames= as (F, natoms, 3) over ALL atoms")
            return np.ascontiguousarray(arr, np.float32)
        from ..io.xtc import open_trajectory
        rd = open_trajectory(dcd)
        out = np.empty((len(rd), natoms, 3), np.float32)
        for f in range(len(rd)):
            out[f] = rd.frame(f)
        return out

    def _scan_dynamic_selections(self, topo, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text):
        """Selections containing ``around`` are re-evaluated on every frame (:319-331): the atoms, their local
        indices and the search grid change per frame.  All frames of every ``around`` are evaluated on the GPU in
        one ``pc_find_within`` call; runs of frames with identical atom sets are then scanned with a topology of
        their own, and their records are re-indexed into the UNION of the per-frame selections, which is what
        the accumulation and the result views work on (results carry global atom indices either way)."""
        from ..io.selection import select_atoms_frames
        from ..scheduler import ArraySource
        self_mode = sel2text == "self"
        dev = (self._devices or [0])[0]
        coords = self._all_frames(dcd, topo.n_atoms)
        F = len(coords)
        per1 = select_atoms_frames(topo, sel1text, coords, device=dev)
        per2 = per1 if self_mode else select_atoms_frames(topo, sel2text, coords, device=dev)
        if F == 0 or len(per1[0]) == 0 or len(per2[0]) == 0:
            raise Exception("empty atom selection")                      # :312-313 (selections of the first frame)
        hcsr = prepare.hydrogen_csr_from_bond_table(topo.names, topo.types, topo.bonds)
        bb = np.zeros(topo.n_atoms, bool)
        bb[self.backbone] = True
        seg_codes = prepare.string_codes(topo.segids)
        prep = lambda idx: prepare.prepare_selection(idx, topo.names, topo.resids, seg_codes, bb, hcsr)
        u1 = np.unique(np.concatenate(per1))
        S1 = prep(u1)
        if self_mode:
            u2, S2 = u1, S1
        else:
            u2 = np.unique(np.concatenate(per2))
            S2 = prep(u2)
        self._sel = (S1, S2)
        self.totalFrameNumber = F
        traj = TrajectoryContacts(S1, S2, hbondcutoff, hbondcutangle)
        stats = dict(frames=F, contacts=0, hbonds=0, pair_evals=0, batches=0, path=0, selection_runs=0)
        start = time.time()
        empty_i, empty_f = np.zeros(0, np.int64), np.zeros(0)
        f = 0
        while f < F:
            g = f + 1
            while g < F and np.array_equal(per1[g], per1[f]) and (self_mode or np.array_equal(per2[g], per2[f])):
                g += 1
            i1, i2 = per1[f], per2[f]
            stats["selection_runs"] += 1
            if len(i1) == 0 or len(i2) == 0:
                part = dict(frame_counts=np.zeros(g - f, np.int64), li=empty_i, lj=empty_i, idx1=empty_i, idx2=empty_i,
                            distance=empty_f, weight=empty_f, hb_count=empty_i, hb_donor=empty_i, hb_acceptor=empty_i,
                            hb_hydrogen=empty_i, hb_side=np.zeros(0, np.int8), hb_distance=empty_f, hb_angle=empty_f)
            else:
                s1 = prep(i1)
                s2 = s1 if self_mode else prep(i2)
                eng = ContactEngine(s1, None if self_mode else s2, cutoff, hbondcutoff, hbondcutangle, device=dev)
                try:
                    src = ArraySource(np.ascontiguousarray(coords[f:g][:, i1]),
                                      None if self_mode else np.ascontiguousarray(coords[f:g][:, i2]))
                    t, _, st = run_scan([eng], src, batch_frames=self._batch_frames, order_keys=True, accumulate=False,
                                        keep_contacts=True)
                finally:
                    eng.close()
                for k in ("contacts", "hbonds", "pair_evals", "batches"):
                    stats[k] += st[k]
                stats["path"] = st["path"]
                part = dict(frame_counts=np.diff(t.frame_ptr), li=np.searchsorted(u1, t.idx1), lj=np.searchsorted(u2, t.idx2),
                            idx1=t.idx1, idx2=t.idx2, distance=t.distance, weight=t.weight, hb_count=np.diff(t.hb_ptr),
                            hb_donor=t.hb_donor, hb_acceptor=t.hb_acceptor, hb_hydrogen=t.hb_hydrogen, hb_side=t.hb_side,
                            hb_distance=t.hb_distance, hb_angle=t.hb_angle)
            traj._parts.append(part)
            traj.frame_numbers.extend(range(f, g))
            f = g
            self.currentFrameNumber = g - 1
            self.frameUpdate.emit(float(g) / float(F))
        traj.finalize()
        stats["seconds"] = time.time() - start
        self.scanStats = stats
        self._traj = traj
        return LazyContactResults(traj)

    # ---- accumulation --------------------------------------------------------------------------------
    def analyze_contactResultsWithMaps(self, contactResults, map1, map2):
        """Group every frame's contacts by the (map1, map2) key on the device and build the
        per-key time series (:502-597)."""
        traj = self._traj_for(contactResults)
        s1, s2 = traj.sel1, traj.sel2
        gm1 = prepare.group_map(s1.indices, map1, self.name_array, self.resid_array, self.resname_array, self.segids)
        gm2 = prepare.group_map(s2.indices, map2, self.name_array, self.resid_array, self.resname_array, self.segids)
        F = traj.n_frames
        table = AccumulationTable(F, gm1, gm2)
        if F:
            self_mode = s1 is s2
            dev = (self._devices or [0])[0]
            # backbone membership may have been replaced through setTrajectoryData
            eng = ContactEngine(_with_backbone(s1, self.backbone), None if self_mode else _with_backbone(s2, self.backbone),
                                self.cutoff, self.hbondcutoff, self.hbondcutangle, device=dev, slots=1)
            try:
                eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
                fp, hp = traj.frame_ptr, traj.hb_ptr
                step = max(1, min(F, int(4e6 // max(1, (fp[-1] // max(F, 1)) + 1))))
                for f0 in range(0, F, step):
                    f1 = min(F, f0 + step)
                    c0, c1 = int(fp[f0]), int(fp[f1])
                    rec, hb, order = _records(traj, c0, c1, fp[f0:f1 + 1])
                    b = eng.accumulate_records(rec, hb, np.diff(fp[f0:f1 + 1]), _hb_per_frame(traj, f0, f1),
                                               order_keys=order, frame0=f0)
                    table.add_batch(b, f0)
                    self.frameUpdate.emit(float(f1) / float(F))
            finally:
                eng.close()
        table.finalize()
        self._table = table
        lists = contactResults if isinstance(contactResults, LazyContactResults) else None
        return table.to_objects(self.name_array, self.resid_array, self.resname_array, self.segids, traj, lists)

    def _traj_for(self, contactResults) -> TrajectoryContacts:
        if isinstance(contactResults, LazyContactResults):
            return contactResults.traj
        return _traj_from_objects(contactResults, self)


# ---- helpers ---------------------------------------------------------------------------------------
def _with_backbone(sel, backbone):
    from .. import _lib
    from ..prepare import SelectionTopology
    flags = sel.flags & ~np.uint8(_lib.PC_ATOM_BACKBONE)
    if len(backbone):
        bb = np.asarray(sorted(int(b) for b in backbone), np.int64)
        pos = np.minimum(np.searchsorted(bb, sel.indices), bb.size - 1)
        flags[bb[pos] == sel.indices] |= np.uint8(_lib.PC_ATOM_BACKBONE)
    return SelectionTopology(sel.indices, flags.astype(np.uint8), sel.resid, sel.segid, sel.hptr, sel.hidx)


def _records(traj: TrajectoryContacts, c0, c1, fptr):
    from .. import _lib
    rec = np.zeros(c1 - c0, _lib.contact_dtype)
    rec["i"], rec["j"] = traj.li[c0:c1], traj.lj[c0:c1]
    rec["distance"], rec["weight"] = traj.distance[c0:c1], traj.weight[c0:c1]
    hp = traj.hb_ptr
    h0, h1 = int(hp[c0]), int(hp[c1])
    hb = np.zeros(h1 - h0, _lib.hbond_dtype)
    owner = np.repeat(np.arange(c0, c1), np.diff(hp[c0:c1 + 1]))
    hb["i"], hb["j"] = traj.li[owner], traj.lj[owner]
    hb["distance"], hb["angle"] = traj.hb_distance[h0:h1], traj.hb_angle[h0:h1]
    # contacts are stored in the reference's order: position in the frame IS the order key
    cnt = np.diff(fptr)
    order = (np.arange(c1 - c0) - np.repeat(fptr[:-1] - c0, cnt)).astype(np.uint64)
    return rec, hb, order


def _hb_per_frame(traj: TrajectoryContacts, f0, f1):
    fp, hp = traj.frame_ptr, traj.hb_ptr
    return np.diff(hp[fp[f0:f1 + 1]])


def _traj_from_objects(contactResults, analyzer: Analyzer) -> TrajectoryContacts:
    """contactResults handed in as real lists of AtomContact (e.g. an imported session,
    core/DataHandler.py:8-18): rebuild the flat arrays over 'selections' made of the atoms
    that occur."""
    from ..prepare import SelectionTopology
    from .. import _lib
    a1 = sorted({int(c.idx1) for fr in contactResults for c in fr})
    a2 = sorted({int(c.idx2) for fr in contactResults for c in fr})
    if not a1:
        a1, a2 = [0], [0]

    def sel(atoms):
        idx = np.asarray(atoms, np.int64)
        n = idx.size
        return SelectionTopology(idx, np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32),
                                 np.zeros(n + 1, np.int32), np.zeros(0, np.int32))
    s1, s2 = sel(a1), sel(a2)
    traj = TrajectoryContacts(s1, s2, analyzer.hbondcutoff, analyzer.hbondcutangle)
    l1 = {g: k for k, g in enumerate(a1)}
    l2 = {g: k for k, g in enumerate(a2)}
    li, lj, dist, w, hbc, counts = [], [], [], [], [], []
    hd, ha, hh, hdist, hang, hside = [], [], [], [], [], []
    for fr in contactResults:
        counts.append(len(fr))
        for c in fr:
            li.append(l1[int(c.idx1)]); lj.append(l2[int(c.idx2)])
            dist.append(c.distance); w.append(c.weight); hbc.append(len(c.hbondinfo))
            for hb in c.hbondinfo:
                hd.append(hb.donorIndex); ha.append(hb.acceptorIndex); hh.append(hb.hydrogenIndex)
                hdist.append(hb.acceptorHydrogenDistance); hang.append(hb.angle)
                hside.append(0 if hb.donorIndex == c.idx1 else 1)
    traj.frame_numbers = list(range(len(contactResults)))
    traj.frame_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    traj.li, traj.lj = np.asarray(li, np.int64), np.asarray(lj, np.int64)
    traj.idx1, traj.idx2 = s1.indices[traj.li] if li else np.zeros(0, np.int64), s2.indices[traj.lj] if lj else np.zeros(0, np.int64)
    traj.distance, traj.weight = np.asarray(dist, float), np.asarray(w, float)
    traj.hb_ptr = np.concatenate([[0], np.cumsum(hbc)]).astype(np.int64)
    traj.hb_donor, traj.hb_acceptor, traj.hb_hydrogen = (np.asarray(v, np.int64) for v in (hd, ha, hh))
    traj.hb_distance, traj.hb_angle = np.asarray(hdist, float), np.asarray(hang, float)
    traj.hb_side = np.asarray(hside, np.int8)
    return traj
