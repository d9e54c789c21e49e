"""Minimal CHARMM/NAMD/X-PLOR PSF topology reader.

The reference obtains topology through ``MDAnalysis.Universe(psf, dcd)``
(PyContact/core/ContactAnalyzer.py:282-300, core/multi_trajectory.py:362-390)
and then reads, per atom: ``resname, resid, name, segid`` and ``atom.bonds``
(each bond's ``type`` = tuple of the two atom *types*, ``indices`` = the two
0-based atom indices).  MDAnalysis is a loader for the hot path, not part of
it; this module supplies exactly those arrays so the path runs without it.

Only the ``!NATOM`` and ``!NBOND`` sections are parsed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class Topology:
    """Per-atom arrays in file order (index == MDAnalysis ``atom.index``)."""
    segids: List[str]
    resids: np.ndarray            # int64
    resnames: List[str]
    names: List[str]
    types: List[str]
    charges: np.ndarray
    masses: np.ndarray
    bonds: np.ndarray             # (nbonds, 2) int64, 0-based, file order
    _bond_csr: tuple = field(default=None, repr=False)

    @property
    def n_atoms(self) -> int:
        return len(self.names)

    def bonds_of_atoms(self):
        """CSR (ptr, bond_id) of the bonds touching each atom, each atom's bonds
        ordered by their (i, j) index tuple -- the order MDAnalysis'
        ``atom.bonds`` yields (it builds the group from ``sorted(unique_bonds)``)."""
        if self._bond_csr is None:
            nb = len(self.bonds)
            order = np.lexsort((self.bonds[:, 1], self.bonds[:, 0])) if nb else np.zeros(0, np.int64)
            sb = self.bonds[order]
            atom = np.concatenate([sb[:, 0], sb[:, 1]])
            bid = np.concatenate([order, order])
            rank = np.concatenate([np.arange(nb), np.arange(nb)])
            o2 = np.lexsort((rank, atom))
            counts = np.bincount(atom, minlength=self.n_atoms)
            ptr = np.zeros(self.n_atoms + 1, np.int64)
            np.cumsum(counts, out=ptr[1:])
            self._bond_csr = (ptr, bid[o2])
        return self._bond_csr


def read_psf(path: str) -> Topology:
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    if not lines or not lines[0].startswith("PSF"):
        raise ValueError("%s: not a PSF file" % path)
    i = 0
    n = len(lines)
    segids, resids, resnames, names, types, charges, masses = [], [], [], [], [], [], []
    bonds = np.zeros((0, 2), np.int64)
    while i < n:
        ln = lines[i]
        if "!NATOM" in ln:
            natom = int(ln.split()[0])
            for k in range(natom):
                f = lines[i + 1 + k].split()
                # serial segid resid resname name type charge mass [...]
                segids.append(f[1])
                # resid may carry an insertion code ("12A"); MDAnalysis keeps the number
                r = f[2]
                try:
                    resids.append(int(r))
                except ValueError:
                    resids.append(int("".join(ch for ch in r if (ch.isdigit() or ch == "-"))))
                resnames.append(f[3])
                names.append(f[4])
                types.append(f[5])
                charges.append(float(f[6]))
                masses.append(float(f[7]))
            i += natom + 1
            continue
        if "!NBOND" in ln:
            nbond = int(ln.split()[0])
            vals = []
            j = i + 1
            while len(vals) < 2 * nbond and j < n:
                vals.extend(lines[j].split())
                j += 1
            arr = np.asarray(vals[: 2 * nbond], dtype=np.int64) - 1
            bonds = arr.reshape(-1, 2)
            i = j
            continue
        i += 1
    if not names:
        raise ValueError("%s: no !NATOM section" % path)
    return Topology(segids, np.asarray(resids, np.int64), resnames, names, types,
                    np.asarray(charges), np.asarray(masses), bonds)
