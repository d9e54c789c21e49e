"""Frame-batch engine: the host side of the CUDA hot path.

``ContactEngine`` owns one GPU's contexts of libpycontact_b200.so and runs batches of
frames through it: H2D copy -> scan (search + filters + score + H-bond test) ->
[order keys] -> [accumulate] -> D2H of the records.  It replaces the reference's
frame-chunking scheduler (``run_load_parallel``'s ``chunks`` + ``Pool.apply_async`` +
``loop_trajectory_grid`` per chunk, PyContact/core/multi_trajectory.py:423-451): chunks
become frame batches, pool workers become batches in flight on CUDA streams (two slots:
while one batch computes, the next is uploaded and the previous is read back).

There is no CPU fallback: every number returned was produced by the CUDA library.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Iterable, Iterator, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import PcError, check, ptr
from .prepare import SelectionTopology


@dataclass
class BatchResult:
    """Records of one frame batch as flat arrays (frames are contiguous ranges)."""
    frame0: int
    nframes: int
    contacts: Optional[np.ndarray]       # _lib.contact_dtype, local indices i / j
    order_keys: Optional[np.ndarray]     # uint64 per contact (reference order key) or None
    hbonds: Optional[np.ndarray]         # _lib.hbond_dtype
    rows: Optional[np.ndarray]           # _lib.row_dtype ((frame, key) cells) or None
    cstart: np.ndarray
    ccount: np.ndarray
    hstart: np.ndarray
    hcount: np.ndarray
    rstart: Optional[np.ndarray]
    rcount: Optional[np.ndarray]
    n_pair_evals: int
    path: int                            # 1 = one CTA per frame, 2 = tiled large-frame path


class _Slot:
    """One batch in flight: a library context, a stream and pinned staging buffers."""

    def __init__(self, eng: "ContactEngine"):
        lib = eng.lib
        self.eng = eng
        self.ctx = ctypes.c_void_p()
        check(lib.pc_create(eng.device, eng.cutoff, eng.hbondcutoff, eng.hbondcutangle, int(eng.self_mode),
                            ctypes.byref(self.ctx)))
        self.stream = ctypes.c_void_p()
        check(lib.pc_stream_create(eng.device, ctypes.byref(self.stream)))
        s1, s2 = eng.sel1, eng.sel2
        check(lib.pc_set_topology(
            self.ctx, s1.n, 0 if eng.self_mode else s2.n,
            ptr(s1.flags), None if eng.self_mode else ptr(s2.flags),
            ptr(s1.resid), ptr(s1.segid),
            None if eng.self_mode else ptr(s2.resid), None if eng.self_mode else ptr(s2.segid),
            ptr(s1.hptr), ptr(s1.hidx),
            None if eng.self_mode else ptr(s2.hptr), None if eng.self_mode else ptr(s2.hidx)))
        self.max_frames = 0
        self.cap = [0, 0, 0]
        self.pin_in = [None, None]
        self.pin_out = {}
        self.pending = None
        self.pending_dev = None

    def close(self):
        lib = self.eng.lib
        if self.ctx:
            lib.pc_stream_sync(self.stream)
            lib.pc_destroy(self.ctx)
            lib.pc_stream_destroy(self.stream)
            self.ctx = None
        self.pin_in = [None, None]
        self.pin_out = {}

    def reserve(self, nframes, cap_c, cap_h, cap_r):
        if nframes > self.max_frames or cap_c > self.cap[0] or cap_h > self.cap[1] or cap_r > self.cap[2]:
            self.max_frames = max(self.max_frames, nframes)
            self.cap = [max(self.cap[0], cap_c), max(self.cap[1], cap_h), max(self.cap[2], cap_r)]
            check(self.eng.lib.pc_reserve(self.ctx, self.max_frames, *self.cap))

    def staging(self, which, nfloats):
        buf = self.pin_in[which]
        if buf is None or buf.array.size < nfloats:
            buf = _lib.PinnedArray((nfloats,), np.float32)
            self.pin_in[which] = buf
        return buf.array

    def out(self, name, count, dtype):
        buf = self.pin_out.get(name)
        if buf is None or buf.array.size < count:
            buf = _lib.PinnedArray((max(int(count * 1.25), 1024),), dtype)
            self.pin_out[name] = buf
        return buf.array


class ContactEngine:
    def __init__(self, sel1: SelectionTopology, sel2: Optional[SelectionTopology], cutoff, hbondcutoff,
                 hbondcutangle, device: int = 0, slots: int = 2):
        self.lib = _lib.load()
        self.device = int(device)
        self.cutoff, self.hbondcutoff, self.hbondcutangle = float(cutoff), float(hbondcutoff), float(hbondcutangle)
        self.self_mode = sel2 is None
        self.sel1 = sel1
        self.sel2 = sel1 if sel2 is None else sel2
        if self.sel1.n == 0 or self.sel2.n == 0:
            raise Exception("empty atom selection")       # core/ContactAnalyzer.py:312-313
        self.n1, self.n2 = self.sel1.n, self.sel2.n
        self.slots = [_Slot(self) for _ in range(max(1, slots))]
        self.maps_set = False
        self.n_groups2 = 0
        # per-frame guesses, grown on PC_ERR_CAPACITY
        heavy = int(min((~(self.sel1.flags & 1).astype(bool)).sum(), (~(self.sel2.flags & 1).astype(bool)).sum()))
        self.guess = [max(256, 2 * heavy), max(64, heavy // 4), 256]
        self.lim = [0, 0, 0, 0]          # hits/frame, hbonds/frame, cells, table slots (0 = automatic)
        self.launches = 0
        self.forced_path = 0
        self.pack_bytes = 24
        self.tail = 3

    # ------------------------------------------------------------------------------------------
    def close(self):
        for s in self.slots:
            s.close()
        self.slots = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_maps(self, gid1: np.ndarray, gid2: np.ndarray, n_groups2: int):
        g1 = np.ascontiguousarray(gid1, np.int32)
        g2 = np.ascontiguousarray(gid2, np.int32)
        if g1.size != self.n1 or g2.size != self.n2:
            raise ValueError("group id arrays must cover the selections")
        for s in self.slots:
            check(self.lib.pc_set_maps(s.ctx, ptr(g1), ptr(g2), int(n_groups2)))
        self.maps_set = True
        self.n_groups2 = int(n_groups2)

    def set_limits(self, hits=0, hbonds=0, cells=0, table=0):
        self.lim = [int(hits), int(hbonds), int(cells), int(table)]
        for s in self.slots:
            check(self.lib.pc_set_limits(s.ctx, *self.lim))

    def fold_sparse_init(self, entries: int):
        """Clear (and size) the engine's sparse aggregated table (key -> sums over frames); all slots fold
        into the table of slot 0."""
        s0 = self.slots[0]
        for s in self.slots[1:]:
            check(self.lib.pc_stream_sync(s.stream))
            check(self.lib.pc_fold_sparse_share(s.ctx, s0.ctx))
        check(self.lib.pc_fold_sparse_init(s0.ctx, int(entries), s0.stream))
        check(self.lib.pc_stream_sync(s0.stream))
        self.fold_entries = int(entries)

    def fold_sparse_fetch(self, sort: bool = True) -> np.ndarray:
        """Occupied entries of the table (_lib.fold_entry_dtype), sorted by key unless sort=False (then in table
        order: a view of the slot's pinned buffer, valid until the next fetch)."""
        for s in self.slots[1:]:
            check(self.lib.pc_stream_sync(s.stream))
        s0 = self.slots[0]
        buf = s0.out("fold", self.fold_entries, _lib.fold_entry_dtype)
        n = ctypes.c_int64()
        check(self.lib.pc_fold_sparse_fetch(s0.ctx, s0.stream, ptr(buf), self.fold_entries, ctypes.byref(n)))
        ent = buf[:n.value]
        if not sort:
            return ent
        # keys are unique; sorting a contiguous copy of the key column and gathering whole entries is 6x faster than
        # argsort on the 40-byte-strided field (0.27 vs 1.6 ms for 13 k entries -- a third of a 1250-frame step)
        return np.take(ent, np.argsort(np.ascontiguousarray(ent["key"])), axis=0)

    def fold_sparse_export(self, d_out, max_entries: int):
        """The table's occupied entries -> d_out[0 .. max_entries) on the device (compacted, padded with empty
        keys): one rank's contribution to the all-gather of a frame-sharded run.  Asynchronous on slot 0's stream;
        the other slots must have finished folding (their streams are synchronised here)."""
        for s in self.slots[1:]:
            check(self.lib.pc_stream_sync(s.stream))
        s0 = self.slots[0]
        check(self.lib.pc_fold_sparse_export(s0.ctx, s0.stream, d_out, int(max_entries)))

    def fold_sparse_import(self, d_in, n_entries: int):
        """Fold another rank's exported buffer into this engine's table (slot 0's stream)."""
        s0 = self.slots[0]
        check(self.lib.pc_fold_sparse_import(s0.ctx, s0.stream, d_in, int(n_entries)))

    def set_pack_format(self, bytes_per_row: int):
        """Transfer format of packed rows: 24 (pc_row_packed) or 12 (pc_row_slim: 32-bit key, score, counts;
        the per-key bb sums then come from the aggregated table of pc_fold_rows)."""
        for s in self.slots:
            check(self.lib.pc_set_pack_format(s.ctx, int(bytes_per_row)))
        self.pack_bytes = int(bytes_per_row)

    def set_tail(self, mode):
        """Large-frame path: 3 / True = fused tail kernel that reads selection 1 from the frame (default), 1 = fused
        tail kernel on the compacted atoms of both selections, 0 / False = multi-kernel sequence."""
        self.tail = 3 if mode is True else int(mode)
        for s in self.slots:
            check(self.lib.pc_set_tail(s.ctx, self.tail))

    def set_dense_mode(self, mode: int):
        """Multi-kernel sequence, tests: 0 = pair kernel chosen per frame, 1 = thread per atom, 2 = warp per cell."""
        for s in self.slots:
            check(self.lib.pc_set_dense_mode(s.ctx, int(mode)))

    def set_group_mode(self, on):
        """Large-frame path without the tail kernel: True (default) = group kernel (search + records + accumulation per
        unit of whole groups) when the maps allow it, False = multi-kernel sequence + global table."""
        for s in self.slots:
            check(self.lib.pc_set_group_mode(s.ctx, int(bool(on))))

    def last_group(self, slot_id=0) -> bool:
        """True when the slot's last scan ran the group kernel."""
        return bool(self.lib.pc_last_group(self.slots[slot_id].ctx))

    def set_path(self, path: int):
        """0 = automatic, 1 = small-frame kernel, 2 = large-frame path (tests cover both)."""
        self.forced_path = int(path)
        for s in self.slots:
            check(self.lib.pc_set_path(s.ctx, int(path)))

    # ------------------------------------------------------------------------------------------
    def _grow(self, slot: _Slot, nframes: int):
        """After PC_ERR_CAPACITY: enlarge whatever overflowed."""
        st = ctypes.c_uint64()
        need = [ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()]
        path = ctypes.c_int32()
        check(self.lib.pc_last_status(slot.ctx, ctypes.byref(st), ctypes.byref(need[0]), ctypes.byref(need[1]),
                                      ctypes.byref(need[2]), ctypes.byref(path)))
        st = st.value
        if st & _lib.ST_HITS_OVERFLOW:
            self.lim[0] = max(8192, 2 * (self.lim[0] or 4096))
        if st & _lib.ST_HB_OVERFLOW:
            self.lim[1] = max(1024, 2 * (self.lim[1] or 256))
        if st & _lib.ST_TAIL_OVERFLOW:
            # a frame's in-grid atoms do not fit the fused per-frame tail kernel: multi-kernel sequence from now on
            import os
            if os.environ.get("PC_DEBUG"):
                print("pycontact_b200: tail kernel handed a batch over:", self.lib.pc_last_error().decode(), flush=True)
            self.set_tail(0)
        if st & _lib.ST_PRECISION:
            # a (frame, key) cell with more contacts than a float32 sum carries within 1e-5: exact global table
            self.lim[3] = 16384
        if st & _lib.ST_TABLE_OVERFLOW:
            # beyond 2048 slots the library switches to compact shared-memory slots (<= 8192), then to its
            # global-memory table
            self.lim[3] = 16384 if (path.value == 2 or (self.lim[3] or 0) >= 8192) else max(2048, 2 * (self.lim[3] or 1024))
        if (st & _lib.ST_STAGE_OVERFLOW) or self.lim[0] > 16384 or self.lim[1] > 4096:
            # more atoms / records per frame than one CTA can stage in shared memory: use the
            # tiled large-frame path from now on
            if self.forced_path == 1:
                raise PcError(_lib.PC_ERR_CAPACITY, "frame does not fit the small-frame path")
            self.lim[0] = self.lim[1] = 0
            self.set_path(2)
        for s in self.slots:
            check(self.lib.pc_set_limits(s.ctx, *self.lim))
        if st & _lib.ST_OUT_OVERFLOW:
            # the large-frame path slices the buffers per frame: size by the worst frame seen
            per = [max(1, -(-n.value // nframes)) for n in need]
            grow = 4 if path.value == 2 else 1.5
            self.guess = [max(g, int(p * grow) + 64) for g, p in zip(self.guess, per)]

    def submit(self, slot_id: int, xyz1, xyz2, *, frame0=0, order_keys=False, accumulate=False,
               pinned=False):
        """Enqueue one batch on a slot (asynchronous).  xyz*: float32 (nf, n, 3) host arrays."""
        slot = self.slots[slot_id]
        xyz1 = np.asarray(xyz1)
        nf = int(xyz1.shape[0])
        if xyz1.dtype != np.float32 or xyz1.shape[1:] != (self.n1, 3):
            raise ValueError("xyz1 must be float32 with shape (nframes, %d, 3)" % self.n1)
        if not self.self_mode:
            xyz2 = np.asarray(xyz2)
            if xyz2.dtype != np.float32 or xyz2.shape != (nf, self.n2, 3):
                raise ValueError("xyz2 must be float32 with shape (%d, %d, 3)" % (nf, self.n2))
        if accumulate and not self.maps_set:
            raise ValueError("set_maps() before accumulating")
        if pinned:
            h1 = np.ascontiguousarray(xyz1)
            h2 = None if self.self_mode else np.ascontiguousarray(xyz2)
        else:
            h1 = slot.staging(0, nf * self.n1 * 3)[: nf * self.n1 * 3]
            h1[:] = xyz1.reshape(-1)
            h2 = None
            if not self.self_mode:
                h2 = slot.staging(1, nf * self.n2 * 3)[: nf * self.n2 * 3]
                h2[:] = xyz2.reshape(-1)
        slot.pending = dict(h1=h1, h2=h2, nf=nf, frame0=frame0, order_keys=bool(order_keys),
                            accumulate=bool(accumulate))
        self._launch(slot)

    def _launch(self, slot: _Slot):
        p = slot.pending
        nf = p["nf"]
        slot.reserve(nf, nf * self.guess[0], nf * self.guess[1], nf * self.guess[2])
        check(self.lib.pc_scan_host(slot.ctx, ptr(p["h1"]), ptr(p["h2"]), nf, _lib.PC_LAYOUT_XYZ,
                                    int(p["order_keys"]), int(p["accumulate"]), slot.stream))
        self.launches += 1

    def collect(self, slot_id: int, *, want_contacts=True, want_hbonds=True, copy=True) -> BatchResult:
        """Wait for the slot's batch and read its records back (retrying after growing any
        buffer that overflowed)."""
        slot = self.slots[slot_id]
        p = slot.pending
        if p is None:
            raise RuntimeError("nothing submitted on this slot")
        lib = self.lib
        nf = p["nf"]
        tot = [ctypes.c_int64() for _ in range(4)]
        for attempt in range(12):
            rc = lib.pc_batch_totals(slot.ctx, slot.stream, *[ctypes.byref(t) for t in tot])
            if rc == _lib.PC_ERR_CAPACITY:
                self._grow(slot, nf)
                self._launch(slot)
                continue
            check(rc)
            break
        else:
            raise PcError(_lib.PC_ERR_CAPACITY, "batch still does not fit after growing the buffers")
        nc, nh, nr, ne = (int(t.value) for t in tot)
        path = ctypes.c_int32()
        check(lib.pc_last_status(slot.ctx, None, None, None, None, ctypes.byref(path)))
        acc = p["accumulate"]
        meta = slot.out("meta", 6 * nf, np.uint32)
        m = [meta[k * nf:(k + 1) * nf] for k in range(6)]
        contacts = slot.out("contacts", nc, _lib.contact_dtype)[:nc] if want_contacts else None
        okeys = slot.out("okeys", nc, np.uint64)[:nc] if (want_contacts and p["order_keys"]) else None
        hbonds = slot.out("hbonds", nh, _lib.hbond_dtype)[:nh] if want_hbonds else None
        rows = slot.out("rows", nr, _lib.row_dtype)[:nr] if acc else None
        check(lib.pc_fetch(slot.ctx, slot.stream, ptr(contacts), ptr(hbonds), ptr(rows), ptr(okeys),
                           ptr(m[0]), ptr(m[1]), ptr(m[2]), ptr(m[3]),
                           ptr(m[4]) if acc else None, ptr(m[5]) if acc else None))
        check(lib.pc_stream_sync(slot.stream))
        slot.pending = None
        cp = (lambda a: None if a is None else a.copy()) if copy else (lambda a: a)
        return BatchResult(p["frame0"], nf, cp(contacts), cp(okeys), cp(hbonds), cp(rows),
                           cp(m[0]), cp(m[1]), cp(m[2]), cp(m[3]), cp(m[4]) if acc else None,
                           cp(m[5]) if acc else None, ne, int(path.value))

    def accumulate_records(self, contacts, hbonds, ccount, hcount, order_keys=None, frame0=0) -> BatchResult:
        """Regroup already-fetched records (dense, frame after frame) under the current maps:
        ``runContactAnalysis`` on stored ``contactResults`` (core/ContactAnalyzer.py:64-70)."""
        if not self.maps_set:
            raise ValueError("set_maps() before accumulating")
        slot = self.slots[0]
        lib = self.lib
        nf = int(len(ccount))
        contacts = np.ascontiguousarray(contacts, _lib.contact_dtype)
        hbonds = np.ascontiguousarray(hbonds, _lib.hbond_dtype)
        ccount = np.ascontiguousarray(ccount, np.uint32)
        hcount = np.ascontiguousarray(hcount, np.uint32)
        cstart = (np.cumsum(ccount, dtype=np.uint64) - ccount).astype(np.uint32)
        hstart = (np.cumsum(hcount, dtype=np.uint64) - hcount).astype(np.uint32)
        ok = None if order_keys is None else np.ascontiguousarray(order_keys, np.uint64)
        rows_guess = max(self.guess[2] * nf, 1024)
        for attempt in range(12):
            slot.reserve(nf, max(contacts.size, 1), max(hbonds.size, 1), rows_guess)
            check(lib.pc_load_records(slot.ctx, ptr(contacts), ptr(ok), ptr(hbonds), nf, ptr(cstart), ptr(ccount),
                                      ptr(hstart), ptr(hcount), slot.stream))
            check(lib.pc_accumulate(slot.ctx, slot.stream))
            self.launches += 1
            tot = [ctypes.c_int64() for _ in range(4)]
            rc = lib.pc_batch_totals(slot.ctx, slot.stream, *[ctypes.byref(t) for t in tot])
            if rc == _lib.PC_ERR_CAPACITY:
                st = ctypes.c_uint64()
                need = ctypes.c_int64()
                check(lib.pc_last_status(slot.ctx, ctypes.byref(st), None, None, ctypes.byref(need), None))
                if st.value & _lib.ST_TABLE_OVERFLOW:
                    self.lim[3] = max(2048, 2 * (self.lim[3] or 1024))
                    for s in self.slots:
                        check(lib.pc_set_limits(s.ctx, *self.lim))
                rows_guess = max(int(need.value * 1.25) + 64, 2 * rows_guess if st.value & _lib.ST_OUT_OVERFLOW else rows_guess)
                continue
            check(rc)
            break
        else:
            raise PcError(_lib.PC_ERR_CAPACITY, "accumulation still does not fit after growing the buffers")
        nr = int(tot[2].value)
        rows = np.zeros(nr, _lib.row_dtype)
        rs, rc_ = np.zeros(nf, np.uint32), np.zeros(nf, np.uint32)
        check(lib.pc_fetch(slot.ctx, slot.stream, None, None, ptr(rows), None, None, None, None, None, ptr(rs), ptr(rc_)))
        check(lib.pc_stream_sync(slot.stream))
        return BatchResult(frame0, nf, None, None, None, rows, cstart, ccount, hstart, hcount, rs, rc_, 0, 1)

    def scan_device(self, d_xyz1, d_xyz2, nframes, *, layout=_lib.PC_LAYOUT_XYZ, pitch1=None, pitch2=None,
                    order_keys=False, accumulate=False, slot_id=0, events=None):
        """Scan frames that are already in HBM (device pointers); returns
        (n_contacts, n_hbonds, n_rows, n_pair_evals) after synchronising the slot's stream.
        ``events`` = (start, end) CUDA events recorded around the scan launch only."""
        slot = self.slots[slot_id]
        lib = self.lib
        nf = int(nframes)
        pitch1 = self.n1 * layout if pitch1 is None else int(pitch1)
        pitch2 = self.n2 * layout if pitch2 is None else int(pitch2)
        tot = [ctypes.c_int64() for _ in range(4)]
        for attempt in range(12):
            slot.reserve(nf, nf * self.guess[0], nf * self.guess[1], nf * self.guess[2])
            if events is not None:
                check(lib.pc_event_record(events[0], slot.stream))
            check(lib.pc_scan_device(slot.ctx, d_xyz1, d_xyz2, nf, int(layout), pitch1, pitch2, int(order_keys),
                                     slot.stream))
            if events is not None:
                check(lib.pc_event_record(events[1], slot.stream))
            if accumulate:
                check(lib.pc_accumulate(slot.ctx, slot.stream))
            self.launches += 1
            rc = lib.pc_batch_totals(slot.ctx, slot.stream, *[ctypes.byref(t) for t in tot])
            if rc == _lib.PC_ERR_CAPACITY:
                self._grow(slot, nf)
                continue
            check(rc)
            return tuple(int(t.value) for t in tot)
        raise PcError(_lib.PC_ERR_CAPACITY, "batch still does not fit after growing the buffers")

    def submit_device(self, slot_id, d_xyz1, d_xyz2, nframes, *, layout=_lib.PC_LAYOUT_XYZ, pitch1=None, pitch2=None,
                      order_keys=False, accumulate=False, events=None, pack_rows=False, fold=None):
        """Asynchronous form of scan_device: enqueue scan (+ accumulation) on the slot's stream.
        pack_rows: also pack the rows for transfer; fold=(d_totals, n_keys): also fold them into the
        dense aggregated table -- both right behind the accumulation, so that nothing has to be
        launched (and wait for free SMs) once the host has seen the batch complete."""
        slot = self.slots[slot_id]
        lib = self.lib
        nf = int(nframes)
        p = dict(d1=d_xyz1, d2=d_xyz2, nf=nf, layout=int(layout),
                 pitch1=self.n1 * layout if pitch1 is None else int(pitch1),
                 pitch2=self.n2 * layout if pitch2 is None else int(pitch2),
                 order_keys=int(order_keys), accumulate=bool(accumulate), events=events, pack_rows=bool(pack_rows),
                 fold=fold)
        slot.pending_dev = p
        slot.reserve(nf, nf * self.guess[0], nf * self.guess[1], nf * self.guess[2])
        if events is not None:
            check(lib.pc_event_record(events[0], slot.stream))
        check(lib.pc_scan_device(slot.ctx, p["d1"], p["d2"], nf, p["layout"], p["pitch1"], p["pitch2"], p["order_keys"],
                                 slot.stream))
        if events is not None:
            check(lib.pc_event_record(events[1], slot.stream))
        if accumulate:
            check(lib.pc_accumulate(slot.ctx, slot.stream))
            if fold == "sparse":
                check(lib.pc_fold_rows_sparse(slot.ctx, slot.stream))
            elif fold is not None:
                check(lib.pc_fold_rows(slot.ctx, fold[0], int(fold[1]), slot.stream))
            if pack_rows:
                check(lib.pc_pack_rows(slot.ctx, slot.stream))
        self.launches += 1

    def collect_device(self, slot_id):
        """Wait for submit_device's batch; returns (n_contacts, n_hbonds, n_rows, n_pair_evals).  A batch that
        overflowed a buffer is re-run (synchronously) after growing it."""
        slot = self.slots[slot_id]
        lib = self.lib
        p = slot.pending_dev
        tot = [ctypes.c_int64() for _ in range(4)]
        for attempt in range(12):
            rc = lib.pc_batch_totals(slot.ctx, slot.stream, *[ctypes.byref(t) for t in tot])
            if rc == _lib.PC_ERR_CAPACITY:
                self._grow(slot, p["nf"])
                self.submit_device(slot_id, p["d1"], p["d2"], p["nf"], layout=p["layout"], pitch1=p["pitch1"],
                                   pitch2=p["pitch2"], order_keys=p["order_keys"], accumulate=p["accumulate"],
                                   events=p["events"], pack_rows=p["pack_rows"], fold=p["fold"])
                continue
            check(rc)
            slot.pending_dev = None
            return tuple(int(t.value) for t in tot)
        raise PcError(_lib.PC_ERR_CAPACITY, "batch still does not fit after growing the buffers")

    def fetch_rows(self, slot_id=0, packed=False):
        """Rows of the slot's last accumulated batch -> (rows, rstart, rcount) host arrays (pinned views).
        packed=True transfers pc_row_packed (24 B, float32 sums, no first-appearance key)."""
        slot = self.slots[slot_id]
        lib = self.lib
        nr = ctypes.c_int64()
        st = ctypes.c_uint64()
        nf = slot.max_frames
        check(lib.pc_last_status(slot.ctx, ctypes.byref(st), None, None, ctypes.byref(nr), None))
        n = min(int(nr.value), slot.cap[2])
        meta = slot.out("meta", 6 * nf, np.uint32)
        if packed:
            if self.pack_bytes == 12:
                rows = slot.out("rows_slim", n, _lib.slim_row_dtype)[:n]
            else:
                rows = slot.out("rows_packed", n, _lib.packed_row_dtype)[:n]
            check(lib.pc_fetch_packed_rows(slot.ctx, slot.stream, ptr(rows), ptr(meta[4 * nf:5 * nf]),
                                           ptr(meta[5 * nf:6 * nf])))
        else:
            rows = slot.out("rows", n, _lib.row_dtype)[:n]
            check(lib.pc_fetch(slot.ctx, slot.stream, None, None, ptr(rows), None, None, None, None, None,
                               ptr(meta[4 * nf:5 * nf]), ptr(meta[5 * nf:6 * nf])))
        check(lib.pc_stream_sync(slot.stream))
        return rows, meta[4 * nf:5 * nf], meta[5 * nf:6 * nf]

    # ------------------------------------------------------------------------------------------
    def scan(self, xyz1, xyz2=None, **kw) -> BatchResult:
        """One batch, synchronously."""
        want = {k: kw.pop(k) for k in ("want_contacts", "want_hbonds") if k in kw}
        self.submit(0, xyz1, xyz2, **kw)
        return self.collect(0, **want)

    def scan_batches(self, batches: Iterable[Tuple[int, np.ndarray, Optional[np.ndarray]]], **kw) -> Iterator[BatchResult]:
        """Pipeline ``(frame0, xyz1, xyz2)`` batches over the slots: batch k+1 is staged and
        uploaded while batch k computes.  Yields results in order."""
        want = {k: kw.pop(k) for k in ("want_contacts", "want_hbonds", "copy") if k in kw}
        ns = len(self.slots)
        inflight = []
        k = 0
        for frame0, a, b in batches:
            sid = k % ns
            if len(inflight) == ns:
                yield self.collect(inflight.pop(0), **want)
            self.submit(sid, a, b, frame0=frame0, **kw)
            inflight.append(sid)
            k += 1
        while inflight:
            yield self.collect(inflight.pop(0), **want)
