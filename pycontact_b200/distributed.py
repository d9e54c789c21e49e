"""Frame-sharded runs over several ranks (one process per GPU, ``torch.distributed``).

The reference parallelises over frames only: ``chunks`` cuts the frame list into
``nproc`` contiguous slices, each worker scans its slice, the per-frame results are
concatenated in slice order (PyContact/core/multi_trajectory.py:423-451).  The same
here with ranks instead of pool workers.  There is no data-path collective -- frames are
independent -- only an end-of-run exchange of results:

* per-frame rows (the score time series) are disjoint in the frame dimension:
  ``gather_rows`` concatenates them in rank order on rank 0;
* time-aggregated maps (sum of scores, bb/sc sums, frames with an H-bond per key --
  what ``mean_score`` / ``hbond_percentage`` / ``bb1..sc2`` are computed from,
  core/Biochemistry.py:150-186, core/ContactAnalyzer.py:588-591) are summed with ONE
  ``all_reduce`` over a key-aligned dense buffer (keys come from the static topology, so
  all ranks agree on them without communication).

Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .scheduler import worker_slices


def frame_slice(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the frames rank ``rank`` analyses -- the reference's chunk rule, made lossless
    (scheduler.worker_slices: every frame belongs to exactly one rank)."""
    return worker_slices(n_frames, world)[rank]


def aggregate_rows(rows: np.ndarray, n_keys: int) -> np.ndarray:
    """Dense [n_keys, 4] float64 table (sum score, sum bb1, sum bb2, frames with an H-bond) of a
    rank's rows -- the host twin of ``pc_fold_rows``."""
    out = np.zeros((n_keys, 4))
    if rows.size:
        k = rows["key"].astype(np.int64)
        np.add.at(out[:, 0], k, rows["score"])
        np.add.at(out[:, 1], k, rows["bb1"])
        np.add.at(out[:, 2], k, rows["bb2"])
        np.add.at(out[:, 3], k, (rows["nhbonds"] > 0).astype(np.float64))
    return out


def all_reduce_totals(totals, device: Optional[str] = None):
    """Sum the aggregated maps over ranks (one all_reduce).  ``totals`` is a numpy array (moved to
    ``device`` for NCCL) or a torch tensor already on the right device; returns the same kind."""
    import torch
    import torch.distributed as dist
    if isinstance(totals, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(totals))
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)
        return t.cpu().numpy()
    dist.all_reduce(totals)
    return totals


def gather_rows(rows: np.ndarray, frames: np.ndarray, dst: int = 0):
    """Concatenate every rank's (rows, global frame index per row) on ``dst`` in rank order
    (``allContacts.extend(rn)`` at core/multi_trajectory.py:447-451).  Returns (rows, frames) on
    ``dst`` and (None, None) elsewhere."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = (rows.tobytes(), np.ascontiguousarray(frames, np.int64).tobytes())
    gathered: List = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None, None
    r = np.concatenate([np.frombuffer(g[0], _lib.row_dtype) for g in gathered])
    f = np.concatenate([np.frombuffer(g[1], np.int64) for g in gathered])
    return r, f
