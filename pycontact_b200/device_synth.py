"""Synthetic trajectories generated on the device (bench.py).

``DeviceTrajectory`` keeps frame 0 and the per-residue motion parameters of a
``synth.SynthSystem`` in HBM and writes any range of frames with ``pc_synth_frames``
(counter-based generator keyed (seed, frame, atom), so every rank can generate its own
frame shard).  The analysis input is *defined* as the float32 values this produces;
parity checks copy sampled frames back and run the oracle on exactly those bytes.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import DeviceBuffer, check


class DeviceTrajectory:
    def __init__(self, system, device=0, atoms=None):
        """atoms: global indices to generate (default: all atoms of the system)."""
        self.system = system
        self.device = int(device)
        self.lib = _lib.load()
        atoms = np.arange(system.topology.n_atoms) if atoms is None else np.asarray(atoms)
        self.atoms = atoms
        self.n = int(atoms.size)
        self.base = DeviceBuffer.from_array(system.coords0[atoms].astype(np.float32), device)
        self.res = DeviceBuffer.from_array(system.residue_of[atoms].astype(np.int32), device)
        self.amp = DeviceBuffer.from_array(system.amp.astype(np.float32), device)
        self.period = DeviceBuffer.from_array(system.period.astype(np.float32), device)
        self.phase = DeviceBuffer.from_array(system.phase.astype(np.float32), device)
        # the generator keys its noise by the position in `atoms`; mix the selection's first atom in
        self.seed = (int(system.seed) * 0x9E3779B1 + int(atoms[0]) if atoms.size else int(system.seed)) & (2 ** 63 - 1)

    def generate(self, out_ptr, frame0, nframes, layout=_lib.PC_LAYOUT_XYZ, pitch=None, stream=None):
        pitch = self.n * layout if pitch is None else int(pitch)
        check(self.lib.pc_synth_frames(self.device, self.base.ptr, self.res.ptr, self.amp.ptr, self.period.ptr,
                                       self.phase.ptr, self.n, float(self.system.sigma), self.seed, int(frame0),
                                       int(nframes), int(layout), pitch, out_ptr, stream))

    def frames_to_host(self, frames) -> np.ndarray:
        out = np.empty((len(frames), self.n, 3), np.float32)
        buf = DeviceBuffer(self.n * 12, self.device)
        try:
            for k, f in enumerate(frames):
                self.generate(buf.ptr, int(f), 1)
                check(self.lib.pc_device_sync(self.device))
                out[k] = buf.download(np.float32, self.n * 3).reshape(self.n, 3)
        finally:
            buf.free()
        return out
