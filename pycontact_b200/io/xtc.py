"""GROMACS XTC trajectory reader (frames decoded by the library's host code, csrc/pc_xtc.h).

Stands in for ``for ts in u.trajectory`` of the reference on XTC input
(PyContact/core/ContactAnalyzer.py:282, :319; core/multi_trajectory.py:362, :401-417; the reference's tests open
``exampleData/md_noPBC.xtc``, tests/test_basic.py:28-30): float32 positions per frame in Angstrom (MDAnalysis scales
the file's nm by 10).  Same interface as :class:`pycontact_b200.io.dcd.DCDReader`, so
:class:`pycontact_b200.scheduler.DCDSource` feeds the scan from either format batch by batch.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib


class XTCReader:
    def __init__(self, path: str):
        self.path = str(path)
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._lib.pc_xtc_open(self.path.encode(), ctypes.byref(h)))
        self._h = h
        n, f = ctypes.c_int32(), ctypes.c_int64()
        _lib.check(self._lib.pc_xtc_info(self._h, ctypes.byref(n), ctypes.byref(f)))
        self.n_atoms, self.n_frames = int(n.value), int(f.value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pc_xtc_close(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return self.n_frames

    def read_into(self, out: np.ndarray, frame0: int, nframes: int, indices=None, *, steps=None, times=None, boxes=None):
        """Fill ``out[k] = positions(frame0 + k)[indices]`` for a frame batch; ``out`` is a C-contiguous float32
        (nframes, n, 3) array, typically a pinned staging buffer."""
        idx = None if indices is None else np.ascontiguousarray(indices, np.int32)
        n = self.n_atoms if idx is None else int(idx.size)
        if out.dtype != np.float32 or not out.flags.c_contiguous or out.shape[0] < nframes or out.shape[1:] != (n, 3):
            raise ValueError("out must be C-contiguous float32 (nframes, %d, 3)" % n)
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(self._lib.pc_xtc_read(self._h, int(frame0), int(nframes), p(idx), 0 if idx is None else int(idx.size),
                                         p(out), p(steps), p(times), p(boxes)))
        return out

    def frame(self, frame: int, indices=None) -> np.ndarray:
        """float32 (n, 3) positions of ``indices`` (all atoms if None), Angstrom."""
        if not 0 <= frame < self.n_frames:
            raise IndexError(frame)
        n = self.n_atoms if indices is None else len(indices)
        return self.read_into(np.empty((1, n, 3), np.float32), frame, 1, indices)[0]

    def header(self, frame: int):
        """(step, time in ps, box 3 x 3 in Angstrom) of one frame."""
        step, time, box = np.zeros(1, np.int32), np.zeros(1, np.float32), np.zeros((1, 9), np.float32)
        self.read_into(np.empty((1, self.n_atoms, 3), np.float32), frame, 1, None, steps=step, times=time, boxes=box)
        return int(step[0]), float(time[0]), box.reshape(3, 3)

    def __iter__(self):
        for f in range(self.n_frames):
            yield self.frame(f)


def open_trajectory(path):
    """Reader for a trajectory file by its extension (.dcd or .xtc)."""
    low = str(path).lower()
    if low.endswith(".dcd"):
        from .dcd import DCDReader
        return DCDReader(path)
    if low.endswith(".xtc"):
        return XTCReader(path)
    raise NotImplementedError("only DCD and XTC trajectories are read natively; pass frames= for other formats")
