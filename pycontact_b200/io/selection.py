"""A small atom-selection language compatible with the subset of MDAnalysis
selections PyContact's users type into the two selection fields
(``Analyzer(..., sel1text, sel2text)``, PyContact/core/ContactAnalyzer.py:302-310).

Grammar (case-sensitive keywords, like MDAnalysis)::

    expr   := term ('or' term)*
    term   := factor ('and' factor)*
    factor := 'not' factor | '(' expr ')' | primary
    primary:= 'all' | 'protein' | 'backbone' | 'nucleic' (unsupported)
            | 'segid' V+ | 'resname' V+ | 'name' V+ | 'type' V+
            | 'resid' R+ | 'resnum' R+ | 'index' R+ | 'bynum' R+
            | 'around' NUMBER factor | 'same' 'residue' 'as' factor
    V      := string with optional '*' / '?' wildcards
    R      := N | N:M | N-M        (inclusive ranges)

``around R sel`` and ``same residue as sel`` need the positions of one frame
(``select_atoms(topo, text, positions=frame)``): ``around`` = atoms NOT in ``sel`` with an atom of
``sel`` closer than R, decided on the GPU by ``pc_find_within`` with the arithmetic of the reference's
own ``find_within`` (cy_modules/src/gridsearch.C:595-826: float32, strict ``<``).  The reference
re-evaluates such selections per frame (core/ContactAnalyzer.py:321-324); without positions they raise
NotImplementedError.  The result is a sorted array of unique 0-based atom indices, like
``AtomGroup.indices`` of an MDAnalysis selection.
"""
from __future__ import annotations

import fnmatch
import re
from typing import List

import numpy as np

# residue names MDAnalysis' ``protein`` keyword recognises (CHARMM, AMBER,
# GROMOS, OPLS variants); ``backbone`` = protein atoms named N, CA, C, O.
PROTEIN_RESNAMES = frozenset("""
ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL
HSD HSE HSP HID HIE HIP HIS1 HIS2 HISA HISB HISH HISD HISE HSD HSE
CYX CYM CYS1 CYS2 CYSH CYZ ASH GLH LYN LYSH ARGN ASPH GLUH ASF ASN1
ACE NME NMA NH2 AIB ORN DAB MSE ALAD CME CALA CARG CASN CASP CCYS CCYX CGLN CGLU
CGLY CHID CHIE CHIP CILE CLEU CLYS CMET CPHE CPRO CSER CTHR CTRP CTYR CVAL
NALA NARG NASN NASP NCYS NCYX NGLN NGLU NGLY NHID NHIE NHIP NILE NLEU NLYS NMET
NPHE NPRO NSER NTHR NTRP NTYR NVAL CLEU CME DALA GLYM HYP NORL PGLU QLN TRPO TYRO
""".split())
BACKBONE_NAMES = frozenset(["N", "CA", "C", "O"])

_KEYWORDS = {"and", "or", "not", "(", ")", "all", "protein", "backbone", "segid", "resname",
             "name", "type", "resid", "resnum", "index", "bynum", "around", "same", "self"}


class SelectionError(Exception):
    pass


def _tokenize(text: str) -> List[str]:
    return re.findall(r"\(|\)|[^\s()]+", text)


class _Parser:
    def __init__(self, text, topo, positions=None, device=0):
        self.tok = _tokenize(text)
        self.pos = 0
        self.topo = topo
        self.n = topo.n_atoms
        if positions is not None:
            positions = np.ascontiguousarray(positions, np.float32)
            if positions.ndim == 2:
                positions = positions[None]
            if positions.ndim != 3 or positions.shape[1:] != (self.n, 3):
                raise ValueError("positions must be (natoms, 3) or (frames, natoms, 3)")
        self.positions = positions
        self.device = device

    def peek(self):
        return self.tok[self.pos] if self.pos < len(self.tok) else None

    def take(self):
        t = self.peek()
        self.pos += 1
        return t

    def values(self):
        vals = []
        while self.peek() is not None and self.peek() not in _KEYWORDS:
            vals.append(self.take())
        if not vals:
            raise SelectionError("selection keyword needs at least one value")
        return vals

    def expr(self):
        m = self.term()
        while self.peek() == "or":
            self.take()
            m = m | self.term()
        return m

    def term(self):
        m = self.factor()
        while self.peek() == "and":
            self.take()
            m = m & self.factor()
        return m

    def factor(self):
        t = self.peek()
        if t == "not":
            self.take()
            return ~self.factor()
        if t == "(":
            self.take()
            m = self.expr()
            if self.take() != ")":
                raise SelectionError("unbalanced parentheses")
            return m
        return self.primary()

    def _match_strings(self, arr, vals):
        arr = np.asarray(arr, dtype=object)
        m = np.zeros(self.n, bool)
        uniq = set(arr.tolist())
        for v in vals:
            if "*" in v or "?" in v:
                hit = {u for u in uniq if fnmatch.fnmatchcase(u, v)}
                m |= np.fromiter((a in hit for a in arr), bool, self.n)
            else:
                m |= (arr == v)
        return m

    def _match_ranges(self, arr, vals):
        m = np.zeros(self.n, bool)
        for v in vals:
            mm = re.fullmatch(r"(-?\d+)(?:[:\-](-?\d+))?", v)
            if not mm:
                raise SelectionError("bad range %r" % v)
            lo = int(mm.group(1))
            hi = int(mm.group(2)) if mm.group(2) is not None else lo
            m |= (arr >= lo) & (arr <= hi)
        return m

    def primary(self):
        t = self.take()
        topo = self.topo
        if t is None:
            raise SelectionError("unexpected end of selection")
        if t == "all":
            return np.ones(self.n, bool)
        if t == "protein":
            return np.fromiter((r in PROTEIN_RESNAMES for r in topo.resnames), bool, self.n)
        if t == "backbone":
            return np.fromiter(((r in PROTEIN_RESNAMES) and (a in BACKBONE_NAMES)
                                for r, a in zip(topo.resnames, topo.names)), bool, self.n)
        if t == "segid":
            return self._match_strings(topo.segids, self.values())
        if t == "resname":
            return self._match_strings(topo.resnames, self.values())
        if t == "name":
            return self._match_strings(topo.names, self.values())
        if t == "type":
            return self._match_strings(topo.types, self.values())
        if t in ("resid", "resnum"):
            return self._match_ranges(np.asarray(topo.resids), self.values())
        if t == "index":
            return self._match_ranges(np.arange(self.n), self.values())
        if t == "bynum":
            return self._match_ranges(np.arange(1, self.n + 1), self.values())
        if t == "around":
            try:
                r = float(self.take())
            except (TypeError, ValueError):
                raise SelectionError("'around' needs a distance")
            inner = self.factor()
            if self.positions is None:
                raise NotImplementedError("'around' is a per-frame geometric selection: pass positions= "
                                          "(one frame, (natoms, 3)) to select_atoms")
            from ..cy_modules.cy_gridsearch import find_within_frames
            pos = self.positions                                   # (F, natoms, 3)
            if inner.ndim == 1:                                    # one call for all frames
                hit = find_within_frames(pos, (~inner).astype(np.int32), inner.astype(np.int32), r, device=self.device)
            else:                                                  # the inner selection changes per frame itself
                hit = np.stack([find_within_frames(pos[f:f + 1], (~inner[f]).astype(np.int32), inner[f].astype(np.int32),
                                                   r, device=self.device)[0] for f in range(len(pos))])
            return hit.astype(bool)
        if t == "same":
            if self.take() != "residue" or self.take() != "as":
                raise SelectionError("only 'same residue as' is supported")
            inner = self.factor()
            seg = np.asarray(topo.segids, dtype=object)
            res = np.asarray(topo.resids)
            # residues = runs of equal (segid, resid); a residue is in when any of its atoms is
            change = np.ones(self.n, bool)
            change[1:] = (seg[1:] != seg[:-1]) | (res[1:] != res[:-1])
            rid = np.cumsum(change) - 1
            nres = int(rid[-1]) + 1 if self.n else 0
            if inner.ndim == 1:
                has = np.zeros(nres, bool)
                has[rid[inner]] = True
                return has[rid]
            out = np.zeros(inner.shape, bool)
            for f in range(inner.shape[0]):
                has = np.zeros(nres, bool)
                has[rid[inner[f]]] = True
                out[f] = has[rid]
            return out
        raise SelectionError("unknown selection token %r" % t)


def _mask(topo, text, positions, device):
    p = _Parser(text, topo, positions, device)
    mask = p.expr()
    if p.peek() is not None:
        raise SelectionError("trailing tokens in selection: %r" % p.tok[p.pos:])
    return mask


def select_atoms(topo, text: str, positions=None, device=0) -> np.ndarray:
    """Indices (sorted, unique, int64) of the atoms of ``topo`` matching ``text``.  ``positions`` (one frame,
    (natoms, 3)) is needed by the geometric keyword ``around`` only."""
    if positions is not None and np.ndim(positions) != 2:
        raise ValueError("select_atoms takes one frame; use select_atoms_frames for (frames, natoms, 3)")
    mask = _mask(topo, text, positions, device)
    if mask.ndim == 2:
        mask = mask[0]
    return np.nonzero(mask)[0].astype(np.int64)


def select_atoms_frames(topo, text: str, frames, device=0) -> List[np.ndarray]:
    """The selection re-evaluated on every frame of ``frames`` ((F, natoms, 3)), as the reference's serial frame
    loop does for selections containing ``around`` (core/ContactAnalyzer.py:321-324): a list of F index arrays.
    All frames of an ``around`` go to the GPU in one ``pc_find_within`` call."""
    frames = np.asarray(frames)
    F = frames.shape[0]
    mask = _mask(topo, text, frames, device) if F else np.zeros(topo.n_atoms, bool)
    if mask.ndim == 1:
        idx = np.nonzero(mask)[0].astype(np.int64)
        return [idx for _ in range(F)]
    return [np.nonzero(mask[f])[0].astype(np.int64) for f in range(F)]
