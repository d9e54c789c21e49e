"""Static topology pre-pass (host, once per analysis).

Everything the reference's Python pair loop re-derives for every raw pair is a function
of the topology only (PyContact/core/multi_trajectory.py:83-155 ==
core/ContactAnalyzer.py:353-429): "is a hydrogen" (name starts with "H", :89), "can take
part in an H-bond" (name[0] in O/N/S, :112), the residue/segment ids of the
self-interaction filter (:93-95), which hydrogens are bonded to an atom (:115-155), and
the accumulation key of an atom under a map (core/ContactAnalyzer.py:107-158).  This
module turns those into flat arrays for ``pc_set_topology`` / ``pc_set_maps``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

MAP_TOKENS = ("i.", "nm.", "r.", "rn.", "s.")     # AccumulationMapIndex.mapping, core/Biochemistry.py:338
HBOND_FIRST_LETTERS = ("O", "N", "S")              # HydrogenBondAtoms.atoms, core/Biochemistry.py:32-33


def _starts_with_h(strings) -> np.ndarray:
    return np.fromiter((s[:1] == "H" for s in strings), bool, len(strings))


def hydrogen_csr_from_bond_table(names: Sequence[str], types: Sequence[str], bonds: np.ndarray):
    """CSR (ptr, hyd) over ALL atoms of the hydrogens "bound to" each atom under the reference's
    rule (core/multi_trajectory.py:125-139): walk the atom's bonds in
    MDAnalysis order (sorted by index pair); a bond counts when either atom TYPE starts with
    "H"; its hydrogen is the first atom of the index pair whose NAME starts with "H".
    """
    n = len(names)
    bonds = np.asarray(bonds, np.int64).reshape(-1, 2)
    if len(bonds) == 0:
        return np.zeros(n + 1, np.int64), np.zeros(0, np.int64)
    th, nh = _starts_with_h(types), _starts_with_h(names)
    a, b = bonds[:, 0], bonds[:, 1]
    hyd = np.where(nh[a], a, np.where(nh[b], b, -1))
    ok = (th[a] | th[b]) & (hyd >= 0)
    a, b, hyd = a[ok], b[ok], hyd[ok]
    # each qualifying bond contributes its hydrogen to both of its atoms; per atom the bonds
    # stay in (i, j) order
    owner = np.concatenate([a, b])
    first = np.concatenate([a, a])
    second = np.concatenate([b, b])
    hh = np.concatenate([hyd, hyd])
    order = np.lexsort((second, first, owner))
    ptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(owner, minlength=n), out=ptr[1:])
    return ptr, hh[order]


def hydrogen_csr_from_bond_objects(bond_objects, names: Sequence[str]):
    """Same CSR from the reference's ``suppl[0]``: one object per global atom with ``.types``
    (a list of atom-type tuples) and ``.to_indices()`` (a list of index pairs) -- the shape of
    ``ConvBond`` (core/multi_trajectory.py:33-47) and of an MDAnalysis ``atom.bonds`` group."""
    n = len(bond_objects)
    nh = _starts_with_h(names)
    ptr = np.zeros(n + 1, np.int64)
    out: List[int] = []
    for g, bo in enumerate(bond_objects):
        types = bo.types() if callable(getattr(bo, "types", None)) else bo.types
        pairs = bo.to_indices()
        for k, tp in enumerate(types):
            if any(str(t)[:1] == "H" for t in tp):
                for j in pairs[k]:
                    if nh[int(j)]:
                        out.append(int(j))
                        break
        ptr[g + 1] = len(out)
    return ptr, np.asarray(out, np.int64)


@dataclass
class SelectionTopology:
    """Device-ready description of one selection (local index = position in the selection)."""
    indices: np.ndarray          # int64 (n,) global atom index of every selection atom
    flags: np.ndarray            # uint8  PC_ATOM_* bits
    resid: np.ndarray            # int32
    segid: np.ndarray            # int32 codes (equal code <=> equal segid string)
    hptr: np.ndarray             # int32 (n+1,)
    hidx: np.ndarray             # int32 LOCAL index of each bonded hydrogen, -1 = not in this selection

    @property
    def n(self) -> int:
        return int(self.indices.size)


def prepare_selection(indices, names, resids, segid_codes, backbone_mask, hcsr) -> SelectionTopology:
    indices = np.ascontiguousarray(indices, np.int64)
    n = indices.size
    nm = [names[int(g)] for g in indices]
    flags = np.zeros(n, np.uint8)
    flags[_starts_with_h(nm)] |= _lib.PC_ATOM_HYDROGEN
    flags[np.fromiter((s[:1] in HBOND_FIRST_LETTERS for s in nm), bool, n)] |= _lib.PC_ATOM_HBCLASS
    if backbone_mask is not None:
        flags[np.asarray(backbone_mask, bool)[indices]] |= _lib.PC_ATOM_BACKBONE
    resid = (np.asarray(resids)[indices].astype(np.int32) if resids is not None and len(resids)
             else np.zeros(n, np.int32))
    segid = (np.asarray(segid_codes)[indices].astype(np.int32) if segid_codes is not None and len(segid_codes)
             else np.zeros(n, np.int32))
    gptr, ghyd = hcsr
    cnt = (gptr[indices + 1] - gptr[indices]).astype(np.int64)
    hptr = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=hptr[1:])
    total = int(hptr[-1])
    if total:
        src = np.repeat(gptr[indices], cnt) + (np.arange(total) - np.repeat(hptr[:-1], cnt))
        gh = ghyd[src]
        # the hydrogen is looked up in the SAME selection as its heavy atom (:158 / :184)
        pos = np.searchsorted(indices, gh) if _is_sorted(indices) else None
        if pos is not None:
            pos = np.minimum(pos, n - 1)
            local = np.where(indices[pos] == gh, pos, -1)
        else:
            lut = {int(g): k for k, g in enumerate(indices)}
            local = np.fromiter((lut.get(int(h), -1) for h in gh), np.int64, total)
    else:
        local = np.zeros(0, np.int64)
    return SelectionTopology(indices, flags, resid, segid, hptr.astype(np.int32), local.astype(np.int32))


def _is_sorted(a) -> bool:
    return bool(a.size < 2 or np.all(a[1:] > a[:-1]))


def string_codes(values) -> np.ndarray:
    """int32 code per element, equal strings <=> equal codes."""
    arr = np.asarray(list(values), dtype=object)
    if arr.size == 0:
        return np.zeros(0, np.int32)
    _, inv = np.unique(arr.astype(str), return_inverse=True)
    return inv.astype(np.int32)


@dataclass
class GroupMap:
    """Accumulation groups of one selection under one map (makeKeyArraysFromMaps,
    core/ContactAnalyzer.py:107-158): atoms with equal selected properties share a group."""
    gid: np.ndarray              # int32 (n,) group id of every selection atom
    n_groups: int
    rep: np.ndarray              # int64 (n_groups,) a representative GLOBAL atom index per group
    amap: tuple

    def key_array(self, g: int, names, resids, resnames, segids) -> List[str]:
        """The key array of group g AS STRINGS -- the reference round-trips every key through
        makeKeyArraysFromKey (core/ContactAnalyzer.py:160-222), App. A.6 Q7."""
        a = int(self.rep[g])
        props = (a, names[a], resids[a], resnames[a], segids[a])
        return [str(props[k]) if self.amap[k] else "none" for k in range(5)]

    def key_half(self, g: int, names, resids, resnames, segids) -> str:
        """This group's half of makeKeyFromKeyArrays (core/ContactAnalyzer.py:251-274)."""
        ka = self.key_array(g, names, resids, resnames, segids)
        return "".join(tok + v for tok, v in zip(MAP_TOKENS, ka) if v != "none")


def group_map(indices, amap, names, resids, resnames, segids) -> GroupMap:
    amap = tuple(1 if bool(v) else 0 for v in amap)
    if len(amap) != 5:
        raise ValueError("accumulation maps have 5 entries (index, name, resid, resname, segid)")
    indices = np.asarray(indices, np.int64)
    n = indices.size
    cols = []
    if amap[0]:
        cols.append(indices)
    if amap[1]:
        cols.append(string_codes([names[int(g)] for g in indices]).astype(np.int64))
    if amap[2]:
        # the key uses str(resid): equal strings <=> equal integers
        cols.append(np.asarray(resids)[indices].astype(np.int64))
    if amap[3]:
        cols.append(string_codes([resnames[int(g)] for g in indices]).astype(np.int64))
    if amap[4]:
        cols.append(string_codes([segids[int(g)] for g in indices]).astype(np.int64))
    if not cols:
        return GroupMap(np.zeros(n, np.int32), 1, indices[:1].copy(), amap)
    order = np.lexsort(tuple(cols[::-1]))
    stacked = np.stack([c[order] for c in cols], 1)
    new = np.ones(n, bool)
    new[1:] = np.any(stacked[1:] != stacked[:-1], axis=1)
    gsorted = np.cumsum(new) - 1
    gid = np.empty(n, np.int32)
    gid[order] = gsorted
    ng = int(gsorted[-1]) + 1
    rep = np.empty(ng, np.int64)
    rep[gsorted[::-1]] = indices[order][::-1]          # smallest position in sort order wins
    return GroupMap(gid, ng, rep, amap)
