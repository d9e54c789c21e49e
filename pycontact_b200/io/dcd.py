"""CHARMM/NAMD DCD trajectory reader (memory-mapped, zero-copy per axis).

Stands in for ``for ts in u.trajectory`` of the reference
(PyContact/core/ContactAnalyzer.py:319, core/multi_trajectory.py:401-417):
yields float32 coordinates per frame.  Unit-cell records are skipped -- the
reference ignores periodic boundaries everywhere.
"""
from __future__ import annotations

import numpy as np


class DCDReader:
    def __init__(self, path: str):
        self.path = path
        raw = np.memmap(path, dtype=np.uint8, mode="r")
        self._raw = raw
        endian = "<"
        first = int(np.frombuffer(raw[:4].tobytes(), "<i4")[0])
        if first != 84:
            if int(np.frombuffer(raw[:4].tobytes(), ">i4")[0]) == 84:
                endian = ">"
            else:
                raise ValueError("%s: not a 32-bit-marker DCD file" % path)
        self._endian = endian
        i4 = endian + "i4"
        if raw[4:8].tobytes() != b"CORD":
            raise ValueError("%s: missing CORD magic" % path)
        icntrl = np.frombuffer(raw[8:88].tobytes(), i4)
        self.n_frames_header = int(icntrl[0])
        self.has_cell = bool(icntrl[10])
        self.has_4d = bool(icntrl[11])
        if int(icntrl[8]) != 0:
            raise NotImplementedError("DCD files with fixed atoms are not supported")
        pos = 92                      # 4 + 84 + 4
        tlen = int(np.frombuffer(raw[pos:pos + 4].tobytes(), i4)[0])
        pos += 4 + tlen + 4
        if int(np.frombuffer(raw[pos:pos + 4].tobytes(), i4)[0]) != 4:
            raise ValueError("%s: bad NATOM record" % path)
        self.n_atoms = int(np.frombuffer(raw[pos + 4:pos + 8].tobytes(), i4)[0])
        pos += 12
        self._first = pos
        rec = 4 + 4 * self.n_atoms + 4
        self._cell_bytes = (4 + 48 + 4) if self.has_cell else 0
        self._frame_bytes = self._cell_bytes + rec * (4 if self.has_4d else 3)
        self._rec = rec
        avail = (raw.size - pos) // self._frame_bytes
        # some writers leave the header count at 0; trust the file size
        self.n_frames = int(avail) if self.n_frames_header in (0,) else min(self.n_frames_header, int(avail))

    def __len__(self):
        return self.n_frames

    def _axis(self, frame: int, ax: int) -> np.ndarray:
        off = self._first + frame * self._frame_bytes + self._cell_bytes + ax * self._rec + 4
        return np.frombuffer(self._raw[off:off + 4 * self.n_atoms], dtype=self._endian + "f4")

    def frame(self, frame: int, indices=None) -> np.ndarray:
        """float32 (n, 3) positions of ``indices`` (all atoms if None)."""
        if not 0 <= frame < self.n_frames:
            raise IndexError(frame)
        n = self.n_atoms if indices is None else len(indices)
        out = np.empty((n, 3), np.float32)
        for ax in range(3):
            a = self._axis(frame, ax)
            out[:, ax] = a if indices is None else a[indices]
        return out

    def _frames_view(self):
        """The whole file as a (frames, 3, atoms) float32 array: the X, Y, Z records of every frame sit at fixed offsets,
        so this is a strided view of the memory map (None when the records are not 4-byte aligned in the file)."""
        if not hasattr(self, "_v"):
            start = self._first + self._cell_bytes + 4
            self._v = None
            if self.n_frames > 0 and start % 4 == 0 and self._frame_bytes % 4 == 0:
                self._v = np.ndarray((self.n_frames, 3, self.n_atoms), dtype=self._endian + "f4", buffer=self._raw, offset=start,
                                     strides=(self._frame_bytes, self._rec, 4))
        return self._v

    def read_into(self, out: np.ndarray, frame0: int, nframes: int, indices=None):
        """Fill ``out[k] = positions(frame0 + k)[indices]`` for a frame batch;
        ``out`` is float32 (nframes, n, 3), typically a pinned staging buffer.  One gather per run of consecutive atom
        indices over all frames and axes of the batch (a segment selection is one run): per frame and axis -- 6 numpy
        calls of ~12 us each -- the Python overhead was most of a C2 frame's 74 us."""
        if not (0 <= frame0 and frame0 + nframes <= self.n_frames):
            raise IndexError((frame0, nframes))
        v = self._frames_view()
        if v is None:
            for k in range(nframes):
                for ax in range(3):
                    a = self._axis(frame0 + k, ax)
                    out[k, :, ax] = a if indices is None else a[indices]
            return out
        blk = v[frame0:frame0 + nframes]                     # (nframes, 3, atoms)
        if indices is None:
            out[...] = blk.transpose(0, 2, 1)
            return out
        idx = np.asarray(indices, np.int64)
        cuts = np.flatnonzero(np.diff(idx) != 1) + 1         # runs of consecutive indices
        if cuts.size + 1 <= 64:
            lo = 0
            for hi in list(cuts) + [idx.size]:
                out[:, lo:hi, :] = blk[:, :, int(idx[lo]):int(idx[hi - 1]) + 1].transpose(0, 2, 1)
                lo = hi
        else:
            out[...] = blk[:, :, idx].transpose(0, 2, 1)
        return out

    def __iter__(self):
        for f in range(self.n_frames):
            yield self.frame(f)
