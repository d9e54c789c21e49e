"""Frame-chunking scheduler.

The reference splits the frame list into ``nproc`` contiguous chunks (``chunks``,
PyContact/core/multi_trajectory.py:20-30) and hands each to a pool worker
(:434-443).  Here a chunk goes to a GPU (one ``ContactEngine`` per device) and is
streamed through it in frame batches sized to keep the copy engines and SMs busy;
results come back in frame order, like the concatenation at :447-451.
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional, Sequence

import numpy as np

from .engine import ContactEngine
from .results import AccumulationTable, TrajectoryContacts


def chunks(seq, num):
    """Split ``seq`` into ``num`` (almost) equal contiguous slices -- same rule (and the same
    float stepping) as core/multi_trajectory.py:20-30, so frame -> worker assignment matches."""
    n = len(seq)
    step = n / float(num)
    out, pos = [], 0.0
    while pos < n:
        out.append(seq[int(pos):int(pos + step)])
        pos += step
    return out


def worker_slices(n_items: int, num: int):
    """``num`` contiguous [lo, hi) ranges that cover range(n_items) exactly once.

    ``chunks`` keeps the reference's float stepping, which for many (n, num) pairs yields num + 1
    non-empty slices (8 frames over 6 workers: 7 slices).  The reference runs every slice as its
    own pool task and loses nothing (core/multi_trajectory.py:434-451); a fixed set of GPUs / ranks
    cannot, so the boundaries of the first ``num`` chunks are kept and whatever ``chunks`` puts
    beyond them is folded into the last worker's range."""
    if num < 1:
        raise ValueError("need at least one worker")
    parts = [p for p in chunks(range(n_items), num) if len(p)]
    out = []
    for w in range(num):
        if w < len(parts) and w < num - 1:
            out.append((parts[w][0], parts[w][-1] + 1))
        elif w == num - 1 and len(parts) > w:
            out.append((parts[w][0], n_items))
        else:
            out.append((0, 0))
    # consecutive, disjoint, complete
    pos = 0
    for lo, hi in out:
        if hi > lo:
            assert lo == pos, (n_items, num, out)
            pos = hi
    assert pos == n_items, (n_items, num, out)
    return out


def default_batch_frames(n1: int, n2: int, target_bytes: int = 96 << 20) -> int:
    per_frame = max(1, (n1 + n2) * 12)
    return int(max(1, min(4096, target_bytes // per_frame)))


class ArraySource:
    """Frames held in host memory: ndarray (F, n, 3) or a list of F arrays (n, 3)."""

    def __init__(self, frames1, frames2=None):
        self.f1, self.f2 = frames1, frames2
        self.n_frames = len(frames1)

    @staticmethod
    def _take(fr, f0, nf):
        if isinstance(fr, np.ndarray) and fr.ndim == 3:
            return np.ascontiguousarray(fr[f0:f0 + nf], np.float32)
        return np.ascontiguousarray(np.stack([np.asarray(fr[f], np.float32).reshape(-1, 3)
                                              for f in range(f0, f0 + nf)]))

    def read(self, f0, nf):
        return self._take(self.f1, f0, nf), (None if self.f2 is None else self._take(self.f2, f0, nf))


class DCDSource:
    """Selection coordinates gathered from a DCD file (``sel.positions`` per frame,
    core/multi_trajectory.py:401-417) batch by batch -- the trajectory is never held whole."""

    def __init__(self, reader, idx1, idx2=None):
        self.reader, self.idx1, self.idx2 = reader, np.asarray(idx1), None if idx2 is None else np.asarray(idx2)
        self.n_frames = len(reader)

    def read(self, f0, nf):
        a = np.empty((nf, self.idx1.size, 3), np.float32)
        self.reader.read_into(a, f0, nf, self.idx1)
        b = None
        if self.idx2 is not None:
            b = np.empty((nf, self.idx2.size, 3), np.float32)
            self.reader.read_into(b, f0, nf, self.idx2)
        return a, b


def prefetched(gen, depth: int = 2):
    """Run a frame-batch generator in a reader thread, up to ``depth`` batches ahead of its consumer.

    The ingest step in front of the path: reading a batch (file I/O + the gather of the
    selected atoms, numpy code that releases the GIL) overlaps the previous batches' H2D copies, kernels
    and the processing of their results, instead of sitting between ``collect`` and the next ``submit``.
    The reference holds the whole trajectory in RAM before it starts (core/multi_trajectory.py:401-417)."""
    if depth <= 0:
        yield from gen
        return
    q: "queue.Queue" = queue.Queue(maxsize=depth)
    end = object()
    stop = threading.Event()

    def reader():
        try:
            for item in gen:
                while not stop.is_set():
                    try:
                        q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put(end)
        except BaseException as exc:          # re-raised in the consumer
            q.put(exc)

    th = threading.Thread(target=reader, daemon=True, name="pycontact-frame-reader")
    th.start()
    try:
        while True:
            item = q.get()
            if item is end:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()


def run_scan(engines: Sequence[ContactEngine], source, *, batch_frames: Optional[int] = None,
             order_keys: bool = True, accumulate: bool = False, keep_contacts: bool = True,
             group_maps=None, frame_numbers: Optional[Sequence[int]] = None,
             progress: Optional[Callable[[float], None]] = None, prefetch: int = 2):
    """Scan every frame of ``source``.  Returns (TrajectoryContacts | None, AccumulationTable | None, stats)."""
    F = source.n_frames
    eng0 = engines[0]
    if batch_frames is None:
        batch_frames = default_batch_frames(eng0.n1, eng0.n2)
    labels = list(range(F)) if frame_numbers is None else list(frame_numbers)
    slices = [range(lo, hi) for lo, hi in worker_slices(F, len(engines)) if hi > lo]
    traj = TrajectoryContacts(eng0.sel1, eng0.sel2, eng0.hbondcutoff, eng0.hbondcutangle) if keep_contacts else None
    table = AccumulationTable(F, *group_maps) if accumulate else None

    def batches(sl):
        for lo in range(sl[0], sl[-1] + 1, batch_frames):
            nf = min(batch_frames, sl[-1] + 1 - lo)
            a, b = source.read(lo, nf)
            yield lo, a, b

    gens = [eng.scan_batches(prefetched(batches(sl), prefetch), order_keys=order_keys, accumulate=accumulate,
                             want_contacts=keep_contacts, want_hbonds=keep_contacts)
            for eng, sl in zip(engines, slices)]
    # A batch's records are put into the reference's order (TrajectoryContacts.make_part: sorts and gathers over
    # millions of records, numpy code that releases the GIL) by worker threads while the scan goes on; the parts are
    # appended in frame order at the end.  Done inline after the scan, this step was the whole run time of the object
    # API on long trajectories (0.1 us per contact against 0.1 us per FRAME on the device).
    per_engine: List[list] = [[] for _ in gens]
    stats = dict(frames=F, contacts=0, hbonds=0, pair_evals=0, batches=0, path=0)
    pool = None                                # created with the first big batch
    alive = list(range(len(gens)))
    done = 0
    try:
        while alive:
            for e in list(alive):
                try:
                    r = next(gens[e])
                except StopIteration:
                    alive.remove(e)
                    continue
                stats["contacts"] += int(r.ccount.sum())
                stats["hbonds"] += int(r.hcount.sum())
                stats["pair_evals"] += r.n_pair_evals
                stats["batches"] += 1
                stats["path"] = r.path
                if table is not None:
                    table.add_batch(r, r.frame0)
                if traj is not None:
                    # small batches are ordered inline at the end: their numpy calls are too short to release the GIL for
                    # long, and worker threads would only take it away from the scan loop (5 ms per hand-over)
                    big = int(r.ccount.sum()) >= 200000
                    if big and pool is None:
                        pool = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) - 1)))
                    per_engine[e].append((pool.submit(traj.make_part, r) if big else r, r.frame0, r.nframes))
                done += r.nframes
                if progress is not None:
                    progress(done / float(F))
        for res in per_engine:                    # engine order == frame order (contiguous slices)
            for item, frame0, nf in res:
                part = item.result() if hasattr(item, "result") else traj.make_part(item)
                traj.append_part(part, labels[frame0:frame0 + nf])
    finally:
        if pool is not None:
            pool.shutdown(wait=True, cancel_futures=True)
    if traj is not None:
        traj.finalize()
    if table is not None:
        table.finalize()
    return traj, table, stats
