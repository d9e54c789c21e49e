"""Solvent-accessible surface over a trajectory -- the computation behind the reference's SASA widget
(PyContact/gui/SasaWidgets.py:27-49 ``calculate_sasa_parallel`` and :136-271 ``calculateSasa``) without the Qt
parts: all frames of a selection go to the GPU in one ``pc_sasa`` call instead of one ``cy_sasa`` call per
frame in a process pool.

Defaults are the widget's hard-wired ones (:156-165): probe radius 1.4, 50 random sphere points
(pointstyle 1), pair distance 2 * (2.0 + 1.4); radii from ``vdwRadius(name[0])``.
"""
from __future__ import annotations

import numpy as np

from .core.Biochemistry import vdwRadius
from .core.multi_trajectory import chunks
from .cy_modules import cy_gridsearch

PROBE_RADIUS = 1.4
SURFACE_POINTS = 50
POINTSTYLE = 1
PAIRDIST = 2.0 * (2.0 + 1.4)


def calculate_sasa_parallel(input_coords, natoms, pairdist, nprad, surfacePoints, probeRadius, pointstyle,
                            restricted, restrictedList, rank, device=0):
    """Same arguments and return value as the reference's worker (gui/SasaWidgets.py:27-49): a list with one
    float per frame of ``input_coords`` (a sequence of (natoms, 3) arrays)."""
    if len(input_coords) == 0:
        return []
    natoms = int(natoms)
    coords = np.empty((len(input_coords), natoms, 3), np.float32)
    for i, c in enumerate(input_coords):
        coords[i] = np.reshape(c, (natoms, 3))
    tot = cy_gridsearch.sasa_frames(coords, nprad, pairdist, surfacePoints, probeRadius, pointstyle, restricted,
                                    restrictedList, device=device)
    return [float(t) for t in tot]


def radii_and_restriction(names, sel_indices, restriction_indices=None):
    """nprad (float32) and restrictedList (int32) for a selection, as calculateSasa builds them (:176-191):
    without a restriction the list is ``[0]`` and ``restricted`` = 0."""
    nprad = np.array([vdwRadius(names[i][0]) for i in sel_indices], dtype=np.float32)
    if restriction_indices is None:
        return nprad, np.array([0], dtype=np.int32), 0
    rs = set(int(i) for i in restriction_indices)
    return nprad, np.array([1 if int(i) in rs else 0 for i in sel_indices], dtype=np.int32), 1


def calculate_sasa(psf, dcd, seltext, resseltext="", seltext2="", contact_area=False, nprocs=1, devices=None,
                   topology=None, frames=None):
    """SASA per frame of ``seltext`` (restricted to ``resseltext`` when given); with ``contact_area`` the
    difference SASA(sel1) - SASA(sel2) per frame, both under the restriction (:221-268).  ``nprocs`` chunks of
    frames are spread over ``devices`` (default: GPU 0).  ``topology`` / ``frames`` ((F, natoms, 3)) replace the
    files for callers that already hold them."""
    from .io.selection import select_atoms
    if topology is None:
        from .io.psf import read_psf
        topology = read_psf(psf)
    if frames is None:
        from .io.xtc import open_trajectory
        rd = open_trajectory(dcd)
        frames = np.stack([rd.frame(f) for f in range(len(rd))]) if len(rd) else np.zeros((0, topology.n_atoms, 3), np.float32)
    frames = np.asarray(frames, np.float32)
    first = frames[0] if len(frames) else None
    devices = list(devices) if devices else [0]

    def run(text):
        sel = select_atoms(topology, text, positions=first)
        # the restriction is evaluated once, like the reference's (the widget's own TODO, :193-194)
        res = select_atoms(topology, resseltext, positions=first) if resseltext != "" else None
        nprad, rl, restricted = radii_and_restriction(topology.names, sel, res)
        coords = [frames[f][sel] for f in range(len(frames))]
        out = []
        for rank, chunk in enumerate(chunks(coords, max(1, int(nprocs)))):
            out.extend(calculate_sasa_parallel(chunk, len(sel), PAIRDIST, nprad, SURFACE_POINTS, PROBE_RADIUS,
                                               POINTSTYLE, restricted, rl, rank, device=devices[rank % len(devices)]))
        return out

    sasas = run(seltext)
    if contact_area:
        if resseltext == "":
            raise ValueError("You need a restricted selection for contact areas!")      # :241
        sasas = [a - b for a, b in zip(sasas, run(seltext2))]
    return sasas
