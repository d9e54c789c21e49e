"""Per-contact statistics on the GPU.

``key_statistics`` evaluates, for every accumulated contact at once, what the reference computes one
object and one Python loop at a time: ``mean_score`` / ``median_score`` / ``hbond_percentage`` /
``total_time`` / ``mean_life_time`` / ``median_life_time`` (PyContact/core/Biochemistry.py:150-213).
Input is the (keys x frames) score matrix (``scoreArray`` of each contact) and the H-bonds per
(key, frame); one CUDA launch (``pc_key_stats``, one CTA per key) returns a (keys x 8) table.
There is no CPU fallback: without the CUDA library the call raises.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import DeviceBuffer, check

COLUMNS = ("mean", "median", "hbond_percentage", "total_time", "mean_life_time", "median_life_time",
           "life_entries", "frames_above")
MAX_FRAMES = 16384


def key_statistics(score_matrix, hbond_matrix=None, ns_per_frame: float = 1.0, threshold: float = 0.0,
                   device: int = 0) -> Dict[str, np.ndarray]:
    """``score_matrix`` [K, F] float64, ``hbond_matrix`` [K, F] integer or None -> dict of [K] float64."""
    lib = _lib.load()
    sc = np.ascontiguousarray(score_matrix, dtype=np.float64)
    if sc.ndim != 2:
        raise ValueError("score_matrix must be (keys, frames)")
    K, F = sc.shape
    if F > MAX_FRAMES:
        raise ValueError("pc_key_stats handles up to %d frames per call" % MAX_FRAMES)
    out = np.zeros((K, len(COLUMNS)))
    if K == 0:
        return {c: out[:, i].copy() for i, c in enumerate(COLUMNS)}
    hb = None
    if hbond_matrix is not None:
        hb = np.ascontiguousarray(hbond_matrix, dtype=np.uint32)
        if hb.shape != sc.shape:
            raise ValueError("hbond_matrix must have the shape of score_matrix")
    d_sc = DeviceBuffer.from_array(sc, device)
    d_hb = DeviceBuffer.from_array(hb, device) if hb is not None else None
    d_out = DeviceBuffer(out.nbytes, device)
    try:
        check(lib.pc_key_stats(device, d_sc.ptr, d_hb.ptr if d_hb is not None else None, K, F, float(ns_per_frame),
                               float(threshold), d_out.ptr, None))
        out = d_out.download(np.float64, K * len(COLUMNS)).reshape(K, len(COLUMNS))
    finally:
        for b in (d_sc, d_out, d_hb):
            if b is not None:
                b.free()
    return {c: out[:, i].copy() for i, c in enumerate(COLUMNS)}


def key_statistics_rows(key_ptr, row_frames, row_scores, row_hbonds=None, n_frames: int = 0, ns_per_frame: float = 1.0,
                        threshold: float = 0.0, device: int = 0) -> Dict[str, np.ndarray]:
    """The same table from the (frame, key) cells themselves, grouped by key: cells ``key_ptr[k]:key_ptr[k+1]`` belong
    to key k (``AccumulationTable.key_ptr / row_frames / rows["score"] / rows["nhbonds"]``).  The dense (keys x frames)
    matrix is never built: every key's row is assembled in shared memory by ``pc_key_stats_rows``."""
    lib = _lib.load()
    kp = np.ascontiguousarray(key_ptr, dtype=np.int64)
    K = kp.size - 1
    F = int(n_frames)
    if F > MAX_FRAMES:
        raise ValueError("pc_key_stats_rows handles up to %d frames per call" % MAX_FRAMES)
    out = np.zeros((max(K, 0), len(COLUMNS)))
    if K <= 0:
        return {c: out[:, i].copy() for i, c in enumerate(COLUMNS)}
    fr = np.ascontiguousarray(row_frames, dtype=np.int32)
    sc = np.ascontiguousarray(row_scores, dtype=np.float64)
    if fr.size != sc.size or int(kp[-1]) != sc.size:
        raise ValueError("key_ptr, row_frames and row_scores do not describe the same cells")
    hb = None if row_hbonds is None else np.ascontiguousarray(row_hbonds, dtype=np.uint32)
    bufs = [DeviceBuffer.from_array(kp, device), DeviceBuffer.from_array(fr if fr.size else np.zeros(1, np.int32), device),
            DeviceBuffer.from_array(sc if sc.size else np.zeros(1), device),
            DeviceBuffer.from_array(hb if hb.size else np.zeros(1, np.uint32), device) if hb is not None else None,
            DeviceBuffer(out.nbytes, device)]
    try:
        check(lib.pc_key_stats_rows(device, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr if bufs[3] is not None else None,
                                    K, F, float(ns_per_frame), float(threshold), bufs[4].ptr, None))
        out = bufs[4].download(np.float64, K * len(COLUMNS)).reshape(K, len(COLUMNS))
    finally:
        for b in bufs:
            if b is not None:
                b.free()
    return {c: out[:, i].copy() for i, c in enumerate(COLUMNS)}


def contact_statistics(contacts: Sequence, ns_per_frame: float = 1.0, threshold: float = 0.0, device: int = 0,
                       assign: bool = True) -> Dict[str, np.ndarray]:
    """Statistics of a list of AccumulatedContact (any object with ``scoreArray`` and ``hbondFramesScan``).

    With ``assign`` the values are also stored where the reference's methods leave them:
    ``meanScore``, ``medianScore``, ``ttime``, ``meanLifeTime``, ``medianLifeTime``."""
    if len(contacts) == 0:
        return {c: np.zeros(0) for c in COLUMNS}
    sc = np.asarray([c.scoreArray for c in contacts], dtype=np.float64)
    hb = np.asarray([c.hbondFramesScan() for c in contacts], dtype=np.int64)
    if hb.shape != sc.shape:
        hb = None
    st = key_statistics(sc, hb, ns_per_frame, threshold, device)
    if assign:
        for i, c in enumerate(contacts):
            c.meanScore = float(st["mean"][i])
            c.medianScore = float(st["median"][i])
            c.ttime = float(st["total_time"][i])
            c.meanLifeTime = float(st["mean_life_time"][i])
            c.medianLifeTime = float(st["median_life_time"][i])
    return st
