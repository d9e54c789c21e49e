"""Array-backed results of the GPU path and their views as reference objects.

The reference materialises one Python object per contact
(``contactResults``: list[F] of list[AtomContact], PyContact/core/ContactAnalyzer.py:488-493)
and accumulates through dictionaries of objects (:502-597).  At GPU rates that is
impossible, so records stay in flat numpy arrays (as fetched from the device) and the
reference's object shapes are produced on access:

* ``TrajectoryContacts``  -- all contacts / H-bonds of a frame scan; ``LazyContactResults``
  is its ``contactResults`` view (``len`` = frames, ``[f]`` = list of ``AtomContact``).
* ``AccumulationTable``   -- the (frame, key) cells computed on the device, merged over
  batches (and ranks), ordered like the reference (first appearance), exposed as
  ``AccumulatedContact`` objects with lazy ``scoreArray`` / ``contributingAtoms``.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import List, Optional

import numpy as np

from .core.Biochemistry import AccumulatedContact, AtomContact, HydrogenBond
from .engine import BatchResult
from .prepare import GroupMap, SelectionTopology


def _csr_position(sel: SelectionTopology):
    """sorted composite keys (owner * n + hydrogen) and each entry's position in its owner's list"""
    n = sel.n
    cnt = np.diff(sel.hptr.astype(np.int64))
    owner = np.repeat(np.arange(n, dtype=np.int64), cnt)
    pos = np.arange(owner.size, dtype=np.int64) - np.repeat(sel.hptr[:-1].astype(np.int64), cnt)
    comp = owner * n + sel.hidx.astype(np.int64)
    o = np.argsort(comp, kind="stable")
    return comp[o], pos[o]


class TrajectoryContacts:
    """Contacts and H-bonds of all scanned frames, in the reference's order."""

    def __init__(self, sel1: SelectionTopology, sel2: SelectionTopology, hbondcutoff, hbondcutangle):
        self.sel1, self.sel2 = sel1, sel2
        self.hbondcutoff, self.hbondcutangle = hbondcutoff, hbondcutangle
        self._hpos = None
        self._packable = None              # the selections' sizes fit the packed sort key
        self.packed_sorts = 0              # batches ordered by the one-pass sort (the others: lexsort)
        self._parts = []
        self.frame_numbers: List[int] = []
        # finalised arrays
        self.frame_ptr = np.zeros(1, np.int64)
        self.li = self.lj = self.idx1 = self.idx2 = np.zeros(0, np.int64)
        self.distance = self.weight = np.zeros(0)
        self.hb_ptr = np.zeros(1, np.int64)
        self.hb_donor = self.hb_acceptor = self.hb_hydrogen = np.zeros(0, np.int64)
        self.hb_distance = self.hb_angle = np.zeros(0)
        self.hb_side = np.zeros(0, np.int8)

    @property
    def n_frames(self) -> int:
        return len(self.frame_numbers)

    def add_batch(self, b: BatchResult, frame_numbers):
        """Sort one batch's records into the reference's order and keep them."""
        self.append_part(self.make_part(b), frame_numbers)

    def append_part(self, part, frame_numbers):
        """Keep a batch prepared by make_part; batches are appended in trajectory order."""
        self.packed_sorts += int(part.pop("packed"))
        self._parts.append(part)
        self.frame_numbers.extend(int(f) for f in frame_numbers)

    def make_part(self, b: BatchResult):
        """One batch's records in the reference's order, as the arrays finalize() concatenates.  Does not touch the
        container (apart from two lazily built, idempotent lookup tables), so batches can be prepared by worker threads
        while the scan goes on -- numpy's sorts and gathers release the GIL (scheduler.run_scan)."""
        if b.contacts is None or b.hbonds is None:
            raise ValueError("batch was fetched without its contact / H-bond records")
        nf = b.nframes
        cnt = b.ccount.astype(np.int64)
        frame_of = np.repeat(np.arange(nf, dtype=np.int64), cnt)
        # records of a frame are contiguous, frames need not be in order in the buffer
        src = np.repeat(b.cstart.astype(np.int64), cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        if self._packable is None:
            self._packable = bool(self.sel1.n < (1 << 20) and self.sel2.n < (1 << 20))
        # The reference's order inside a frame is (atom 1, stencil rank, atom 2) -- the order key -- and frames are
        # contiguous here, so this is a segmented sort.  With selections below 2^20 atoms the whole key fits 64 bits --
        # frame | local i | rank | local j -- and ONE argsort does it (numpy's vectorised 64-bit sort: ~10 ns per contact
        # against ~220 ns for the two stable passes of a lexsort, which was three quarters of this function); otherwise lexsort.
        order = None
        if self._packable and nf < (1 << 19):
            ci, cj = b.contacts["i"][src], b.contacts["j"][src]
            rank = np.uint64(0)
            fits = True
            if b.order_keys is not None:
                # the device's order key is local index 1 << 37 | stencil rank << 32 | local index 2 (pc_order.cuh); any
                # other key (a caller's own) is sorted as it is
                okv = b.order_keys[src]
                smp = np.unique(np.linspace(0, max(okv.size - 1, 0), 4096).astype(np.int64)) if okv.size else np.zeros(0, np.int64)
                fits = bool(np.array_equal(okv[smp] >> np.uint64(37), ci[smp].astype(np.uint64)) and
                            np.array_equal(okv[smp] & np.uint64(0xffffffff), cj[smp].astype(np.uint64)))   # (4096 samples)
                rank = (okv >> np.uint64(32)) & np.uint64(31)
            if fits:
                comp = ((frame_of.astype(np.uint64) << np.uint64(45)) | (ci.astype(np.uint64) << np.uint64(25)) |
                        (rank << np.uint64(20)) | cj.astype(np.uint64))
                order = np.argsort(comp)
        packed = order is not None
        if order is not None:
            src = src[order]                             # (frame_of is the leading part of the key: it stays as it is)
            c = np.take(b.contacts, src)
            li, lj = ci[order].astype(np.int64), cj[order].astype(np.int64)
        else:
            c = b.contacts[src]
            if b.order_keys is not None:
                ok = b.order_keys[src]
            else:
                ok = (c["i"].astype(np.uint64) << np.uint64(32)) | c["j"].astype(np.uint64)
            order = np.lexsort((ok, frame_of))
            c, frame_of = c[order], frame_of[order]
            li, lj = c["i"].astype(np.int64), c["j"].astype(np.int64)
        # ---- H-bonds: attach to their contact, order = atom1's hydrogens then atom2's (CSR order)
        hcnt = b.hcount.astype(np.int64)
        hframe = np.repeat(np.arange(nf, dtype=np.int64), hcnt)
        hsrc = np.repeat(b.hstart.astype(np.int64), hcnt) + (np.arange(hcnt.sum()) - np.repeat(np.cumsum(hcnt) - hcnt, hcnt))
        h = b.hbonds[hsrc]
        n2 = self.sel2.n
        ckey = frame_of * (self.sel1.n * n2) + li * n2 + lj
        hkey = hframe * (self.sel1.n * n2) + h["i"].astype(np.int64) * n2 + h["j"].astype(np.int64)
        # contacts are unique per (frame, i, j): locate each H-bond's contact
        co = np.argsort(ckey, kind="stable")
        where = co[np.searchsorted(ckey[co], hkey)] if hkey.size else np.zeros(0, np.int64)
        side = (h["hyd"] >> np.uint32(31)).astype(np.int8)
        hl = (h["hyd"] & np.uint32(0x7fffffff)).astype(np.int64)
        if self._hpos is None:
            self._hpos = (_csr_position(self.sel1), _csr_position(self.sel2))
        pos = np.zeros(hl.size, np.int64)
        for s, sel, heavy in ((0, self.sel1, h["i"]), (1, self.sel2, h["j"])):
            m = side == s
            if m.any():
                comp, p = self._hpos[s]
                q = heavy[m].astype(np.int64) * sel.n + hl[m]
                pos[m] = p[np.searchsorted(comp, q)]
        horder = np.lexsort((pos, side, where))
        h, side, hl, where = h[horder], side[horder], hl[horder], where[horder]
        g1 = self.sel1.indices[li]
        g2 = self.sel2.indices[lj]
        donor_sel1 = side == 0
        hyd_g = np.where(donor_sel1, self.sel1.indices[np.minimum(hl, self.sel1.n - 1)],
                         self.sel2.indices[np.minimum(hl, n2 - 1)])
        part = dict(
            frame_counts=np.bincount(frame_of, minlength=nf).astype(np.int64),
            li=li, lj=lj, idx1=g1, idx2=g2,
            distance=c["distance"].astype(np.float64), weight=c["weight"].astype(np.float64),
            hb_count=np.bincount(where, minlength=c.size).astype(np.int64),
            hb_donor=np.where(donor_sel1, g1[where], g2[where]) if where.size else np.zeros(0, np.int64),
            hb_acceptor=np.where(donor_sel1, g2[where], g1[where]) if where.size else np.zeros(0, np.int64),
            hb_hydrogen=hyd_g.astype(np.int64), hb_side=side,
            hb_distance=h["distance"].astype(np.float64), hb_angle=h["angle"].astype(np.float64),
            packed=packed)
        return part

    def finalize(self):
        if not self._parts:
            return self
        cat = lambda k: np.concatenate([p[k] for p in self._parts])
        prev_c = self.li.size
        counts = np.concatenate([np.diff(self.frame_ptr)] + [p["frame_counts"] for p in self._parts])
        self.frame_ptr = np.zeros(counts.size + 1, np.int64)
        np.cumsum(counts, out=self.frame_ptr[1:])
        for k in ("li", "lj", "idx1", "idx2", "distance", "weight", "hb_donor", "hb_acceptor", "hb_hydrogen",
                  "hb_side", "hb_distance", "hb_angle"):
            setattr(self, k, np.concatenate([getattr(self, k), cat(k)]))
        hbc = np.concatenate([np.diff(self.hb_ptr)[:prev_c], cat("hb_count")])
        self.hb_ptr = np.zeros(hbc.size + 1, np.int64)
        np.cumsum(hbc, out=self.hb_ptr[1:])
        self._parts = []
        return self

    # ---- object views -----------------------------------------------------------------------------
    def frame_objects(self, f: int, frame_label: Optional[int] = None) -> List[AtomContact]:
        lo, hi = int(self.frame_ptr[f]), int(self.frame_ptr[f + 1])
        label = self.frame_numbers[f] if frame_label is None else frame_label
        out = []
        hp = self.hb_ptr
        for c in range(lo, hi):
            hbs = [HydrogenBond(int(self.hb_donor[k]), int(self.hb_acceptor[k]), int(self.hb_hydrogen[k]),
                                float(self.hb_distance[k]), float(self.hb_angle[k]),
                                self.hbondcutoff, self.hbondcutangle)
                   for k in range(int(hp[c]), int(hp[c + 1]))]
            out.append(AtomContact(label, self.distance[c], self.weight[c], self.idx1[c], self.idx2[c], hbs))
        return out


class LazyContactResults(Sequence):
    """``contactResults``: behaves like list[F] of list[AtomContact]; frames are built on access
    (and cached, so repeated reads return the same objects)."""

    def __init__(self, traj: TrajectoryContacts, frame_label=None):
        self.traj = traj
        self._label = frame_label
        self._cache = {}

    def __len__(self):
        return self.traj.n_frames

    def __getitem__(self, f):
        if isinstance(f, slice):
            return [self[k] for k in range(*f.indices(len(self)))]
        if f < 0:
            f += len(self)
        if not 0 <= f < len(self):
            raise IndexError(f)
        if f not in self._cache:
            self._cache[f] = self.traj.frame_objects(f, self._label)
        return self._cache[f]

    def __reduce__(self):
        return (list, ([self[f] for f in range(len(self))],))


class AccumulationTable:
    """(frame, key) cells from the device merged over all batches, in the reference's key order."""

    def __init__(self, n_frames: int, gm1: GroupMap, gm2: GroupMap):
        self.n_frames = n_frames
        self.gm1, self.gm2 = gm1, gm2
        self._rows = []
        self._frames = []

    def add_batch(self, b: BatchResult, frame_pos0: int):
        """frame_pos0 = position of the batch's first frame in the scanned trajectory."""
        if b.rows is None:
            raise ValueError("batch was not accumulated")
        cnt = b.rcount.astype(np.int64)
        src = np.repeat(b.rstart.astype(np.int64), cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        self._rows.append(np.take(b.rows, src))
        self._frames.append(np.repeat(np.arange(b.nframes, dtype=np.int64) + frame_pos0, cnt))

    def add_packed(self, rows: np.ndarray, rstart, rcount, frame_pos0: int):
        """A batch's rows in the transfer format (_lib.packed_row_dtype).  Without first-appearance
        keys the output order of keys falls back to (first frame, key id)."""
        from ._lib import row_dtype
        cnt = np.asarray(rcount, np.int64)
        src = np.repeat(np.asarray(rstart, np.int64), cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        p = rows[src]
        full = np.zeros(p.size, row_dtype)
        for k in ("key", "score", "bb1", "bb2", "ncontacts", "nhbonds"):
            if k in p.dtype.names:          # the 12-byte slim rows carry no per-cell bb sums
                full[k] = p[k]
        full["first"] = p["key"]
        self._rows.append(full)
        self._frames.append(np.repeat(np.arange(cnt.size, dtype=np.int64) + frame_pos0, cnt))

    def add_rows(self, rows: np.ndarray, frames: np.ndarray):
        self._rows.append(rows)
        self._frames.append(np.asarray(frames, np.int64))

    def finalize(self):
        from ._lib import row_dtype
        rows = np.concatenate(self._rows) if self._rows else np.zeros(0, row_dtype)
        frames = np.concatenate(self._frames) if self._frames else np.zeros(0, np.int64)
        # dense key ids, each key's first appearance (smallest frame, then smallest order key in that frame) and the final
        # (output position, frame) order with two argsorts of contiguous 64-bit keys and segment reductions -- np.unique
        # with inverse plus two two-key lexsorts over all rows cost 1 us per row (2 s for a C2-sized trajectory)
        N = rows.size
        kc = np.ascontiguousarray(rows["key"])
        o = np.argsort(kc)
        ks = kc[o]
        newk = np.ones(N, bool)
        if N:
            newk[1:] = ks[1:] != ks[:-1]
        starts = np.flatnonzero(newk)
        keys = ks[starts]
        K = keys.size
        kid_sorted = np.cumsum(newk) - 1
        kid = np.empty(N, np.int64)
        kid[o] = kid_sorted
        if K:
            fs = frames[o]
            minf = np.minimum.reduceat(fs, starts)
            firsts = np.ascontiguousarray(rows["first"])[o]
            firsts[fs != minf[kid_sorted]] = np.iinfo(np.uint64).max
            minfirst = np.minimum.reduceat(firsts, starts)
            appear = np.lexsort((minfirst, minf))                   # keys in first-appearance order
        else:
            appear = np.zeros(0, np.int64)
        rank_of_key = np.empty(K, np.int64)                          # key -> output position
        rank_of_key[appear] = np.arange(K)
        kpos = rank_of_key[kid]
        # rows sorted by (output position, frame)
        if N and int(frames.max()) < (1 << 32) and int(frames.min()) >= 0:
            o2 = np.argsort((kpos.astype(np.uint64) << np.uint64(32)) | frames.astype(np.uint64))
        else:
            o2 = np.lexsort((frames, kpos))
        self.rows = np.take(rows, o2)                # (np.take on a structured array: 4x faster than fancy indexing)
        self.row_frames = frames[o2]
        self.row_key = kpos[o2]
        self.key_ptr = np.zeros(K + 1, np.int64)
        np.cumsum(np.bincount(kpos, minlength=K), out=self.key_ptr[1:])
        out_keys = np.empty(K, np.uint64)
        out_keys[rank_of_key] = keys
        self.keys = out_keys
        G2 = np.uint64(max(1, self.gm2.n_groups))
        self.g1 = (out_keys // G2).astype(np.int64)
        self.g2 = (out_keys % G2).astype(np.int64)
        # per-key totals (core/ContactAnalyzer.py:588-591); sc = score - bb
        seg = self.key_ptr[:-1]
        red = lambda v: np.add.reduceat(v, seg) if K else np.zeros(0)
        self.total_score = red(self.rows["score"])
        self.bb1 = red(self.rows["bb1"])
        self.bb2 = red(self.rows["bb2"])
        self.sc1 = red(self.rows["score"] - self.rows["bb1"])
        self.sc2 = red(self.rows["score"] - self.rows["bb2"])
        self._rows = self._frames = None
        return self

    @property
    def n_keys(self) -> int:
        return int(self.keys.size)

    def _dense(self, k: int, field: str, dtype):
        lo, hi = int(self.key_ptr[k]), int(self.key_ptr[k + 1])
        out = np.zeros(self.n_frames, dtype)
        out[self.row_frames[lo:hi]] = self.rows[field][lo:hi]
        return out

    def score_row(self, k: int) -> np.ndarray:
        return self._dense(k, "score", np.float64)

    def hbond_row(self, k: int) -> np.ndarray:
        return self._dense(k, "nhbonds", np.int64)

    def ncontrib_row(self, k: int) -> np.ndarray:
        return self._dense(k, "ncontacts", np.int64)

    def score_matrix(self) -> np.ndarray:
        m = np.zeros((self.n_keys, self.n_frames))
        m[self.row_key, self.row_frames] = self.rows["score"]
        return m

    def hbond_matrix(self) -> np.ndarray:
        m = np.zeros((self.n_keys, self.n_frames), np.int64)
        m[self.row_key, self.row_frames] = self.rows["nhbonds"]
        return m

    def statistics(self, ns_per_frame: float = 1.0, threshold: float = 0.0, device: int = 0):
        """mean / median score, hbond_percentage, total_time, mean / median life time of every key
        (core/Biochemistry.py:150-213) from the table's cells, on the GPU (stats.key_statistics_rows)."""
        from . import stats
        return stats.key_statistics_rows(self.key_ptr, self.row_frames, self.rows["score"], self.rows["nhbonds"], self.n_frames,
                                         ns_per_frame, threshold, device)

    def hbond_percentage(self) -> np.ndarray:
        """AccumulatedContact.hbond_percentage for every key (core/Biochemistry.py:150-158)."""
        hits = np.bincount(self.row_key, weights=(self.rows["nhbonds"] > 0), minlength=self.n_keys)
        return hits / float(self.n_frames) * 100.0

    def key_arrays(self, k: int, names, resids, resnames, segids):
        return (self.gm1.key_array(int(self.g1[k]), names, resids, resnames, segids),
                self.gm2.key_array(int(self.g2[k]), names, resids, resnames, segids))

    def to_objects(self, names, resids, resnames, segids, traj: Optional[TrajectoryContacts] = None,
                   contact_lists: Optional[LazyContactResults] = None) -> List[AccumulatedContact]:
        """finalAccumulatedContacts (core/ContactAnalyzer.py:579-592)."""
        out = []
        contrib = _ContribIndex(self, traj, contact_lists) if traj is not None else None
        for k in range(self.n_keys):
            k1, k2 = self.key_arrays(k, names, resids, resnames, segids)
            acc = AccumulatedContact(k1, k2)
            acc._score_src = (lambda k=k: _score_list(self, k))
            acc._hb_src = (lambda k=k: self.hbond_row(k))
            if contrib is not None:
                acc._contrib_src = (lambda k=k: contrib.lists(k))
            else:
                acc._contrib_src = (lambda: [[] for _ in range(self.n_frames)])
            acc.bb1, acc.sc1 = float(self.bb1[k]), float(self.sc1[k])
            acc.bb2, acc.sc2 = float(self.bb2[k]), float(self.sc2[k])
            out.append(acc)
        return out


def _score_list(table: AccumulationTable, k: int):
    # absent (key, frame) cells are the integer 0 in the reference (App. A.6 Q9)
    lo, hi = int(table.key_ptr[k]), int(table.key_ptr[k + 1])
    out = [0] * table.n_frames
    for f, s in zip(table.row_frames[lo:hi].tolist(), table.rows["score"][lo:hi].tolist()):
        out[f] = s
    return out


class _ContribIndex:
    """contributingAtoms[f] of a key = the frame's contacts whose atoms fall in the key's groups,
    in frame order (core/ContactAnalyzer.py:532-546)."""

    def __init__(self, table: AccumulationTable, traj: TrajectoryContacts, lists: Optional[LazyContactResults]):
        self.table, self.traj = table, traj
        self.contact_lists = lists if lists is not None else LazyContactResults(traj)
        self._ck = None

    def _contact_keys(self):
        if self._ck is None:
            t = self.table
            G2 = np.uint64(max(1, t.gm2.n_groups))
            self._ck = (t.gm1.gid[self.traj.li].astype(np.uint64) * G2 + t.gm2.gid[self.traj.lj].astype(np.uint64))
        return self._ck

    def lists(self, k: int):
        t = self.table
        ck = self._contact_keys()
        key = t.keys[k]
        out = [[] for _ in range(t.n_frames)]
        lo, hi = int(t.key_ptr[k]), int(t.key_ptr[k + 1])
        fp = self.traj.frame_ptr
        for f in t.row_frames[lo:hi].tolist():
            a, b = int(fp[f]), int(fp[f + 1])
            objs = self.contact_lists[f]
            out[f] = [objs[c - a] for c in (np.nonzero(ck[a:b] == key)[0] + a).tolist()]
        return out
