"""oracle/frame_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's per-frame contact analysis and of its
accumulation stage.  Every function cites the reference lines it follows.

    bonded_hydrogens   core/ContactAnalyzer.py:387-429  ==  core/multi_trajectory.py:115-155
    frame_contacts     core/multi_trajectory.py:65-219  ==  core/ContactAnalyzer.py:326-492
    accumulate         core/ContactAnalyzer.py:502-597 (+107-158, 251-274, 160-222),
                       core/Biochemistry.py:57-158, 283-296

Pinning: tests/test_oracle.py checks these against tests/golden/*.npz, which
tests/golden/make_golden.py produced by running the UNMODIFIED reference
(``loop_trajectory_grid`` + ``Analyzer.analyze_contactResultsWithMaps``, imported from
/root/reference with MDAnalysis/PyQt5 stubbed) and from the reference's own
pickled session ``exampleData/defaultsession_py3``.

The neighbour search itself is delegated to oracle/native.py (C restatement or the
compiled reference).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

from . import native

HBOND_ELEMENTS = ("O", "N", "S")          # core/Biochemistry.py:32-33
MAP_TOKENS = ["i.", "nm.", "r.", "rn.", "s."]   # core/Biochemistry.py:338


def weight_function(d):
    """core/multi_trajectory.py:15-17"""
    return 1.0 / (1.0 + np.exp(5.0 * (d - 4.0)))


def bonded_hydrogens(names, types, bonds):
    """CSR (ptr, hidx) of the hydrogens 'bound to' every atom by the reference's rule
    (SURVEY.md App. A.4): walk the atom's bonds in MDAnalysis order (sorted by index
    tuple); a bond counts when one of its two atom TYPES starts with "H"; the
    hydrogen is the first atom of the bond's index pair whose NAME starts with "H".
    """
    n = len(names)
    bonds = np.asarray(bonds, np.int64).reshape(-1, 2)
    order = np.lexsort((bonds[:, 1], bonds[:, 0])) if len(bonds) else np.zeros(0, np.int64)
    per_atom = [[] for _ in range(n)]
    name_h = [nm.startswith("H") for nm in names]
    type_h = [tp.startswith("H") for tp in types]
    for b in order:
        i, j = int(bonds[b, 0]), int(bonds[b, 1])
        if not (type_h[i] or type_h[j]):
            continue
        h = i if name_h[i] else (j if name_h[j] else -1)
        if h < 0:
            continue
        per_atom[i].append(h)
        per_atom[j].append(h)
    ptr = np.zeros(n + 1, np.int64)
    for a in range(n):
        ptr[a + 1] = ptr[a] + len(per_atom[a])
    hidx = np.fromiter((h for lst in per_atom for h in lst), np.int64, int(ptr[-1]))
    return ptr, hidx


def _norm3(v):
    return np.sqrt(np.einsum("ij,ij->i", v, v))


def frame_contacts(xyz1, xyz2, indices1, indices2, cutoff, hbondcutoff, hbondcutangle,
                   names, hcsr, resids=None, segids=None, self_interaction=False,
                   search=None):
    """One iteration of the reference frame loop (core/multi_trajectory.py:65-219).

    xyz1/xyz2 : float32 (n,3) positions of the two selections
    indices1/2: global atom index of every selection atom
    hcsr      : bonded_hydrogens() over ALL atoms (global indices)
    Returns a dict of arrays; contacts are in the reference's order
    (idx1-ascending, then stencil rank, then idx2 -- App. A.3), H-bonds are grouped per
    contact through ``hb_ptr`` in the order atom1's hydrogens then atom2's.
    """
    if search is None:
        search = native.oracle_find_contacts
    xyz1 = np.ascontiguousarray(xyz1, np.float32).reshape(-1, 3)
    xyz2 = np.ascontiguousarray(xyz2, np.float32).reshape(-1, 3)
    indices1 = np.asarray(indices1, np.int64)
    indices2 = np.asarray(indices2, np.int64)
    n1 = xyz1.shape[0]
    row_ptr, cols = search(xyz1, xyz2, cutoff)                      # :79
    li = np.repeat(np.arange(n1, dtype=np.int64), np.diff(row_ptr))
    lj = cols.astype(np.int64)
    n_raw = li.size
    g1 = indices1[li]
    g2 = indices2[lj]
    is_h = np.fromiter((nm.startswith("H") for nm in names), bool, len(names))   # re.match("H(.*)") :89
    keep = ~(is_h[g1] | is_h[g2])
    if self_interaction:                                            # :93-95 (signed difference, Q2)
        rs = np.asarray(resids, np.int64)
        sg = np.asarray(segids, dtype=object)
        keep &= ~(((rs[g1] - rs[g2]) < 5) & (sg[g1] == sg[g2]))
    li, lj, g1, g2 = li[keep], lj[keep], g1[keep], g2[keep]
    p1 = xyz1.astype(np.float64)
    p2 = xyz2.astype(np.float64)
    dvec = p1[li] - p2[lj]                                          # :98
    dist = np.sqrt(np.einsum("ij,ij->i", dvec, dvec))               # :99
    weight = weight_function(dist)                                  # :107

    # ---- hydrogen bonds (:112-209) ------------------------------------------------
    hb_class = np.fromiter((nm[0] in HBOND_ELEMENTS for nm in names), bool, len(names))
    cand = np.nonzero(hb_class[g1] & hb_class[g2])[0]
    hptr, hidx = hcsr
    loc1 = {int(g): k for k, g in enumerate(indices1)}
    loc2 = loc1 if indices2 is indices1 else {int(g): k for k, g in enumerate(indices2)}

    recs = []   # (contact, side, order, donor, acceptor, hydrogen, dist, angle)
    for side in (1, 2):
        heavy_g = g1[cand] if side == 1 else g2[cand]
        cnt = (hptr[heavy_g + 1] - hptr[heavy_g]).astype(np.int64)
        if cnt.sum() == 0:
            continue
        c_rep = np.repeat(cand, cnt)
        start = np.repeat(hptr[heavy_g], cnt)
        within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        hg = hidx[start + within]
        loc = loc1 if side == 1 else loc2
        try:
            # the hydrogen is looked up in the SAME selection as its heavy atom (:158, :184)
            hl = np.fromiter((loc[int(h)] for h in hg), np.int64, hg.size)
        except KeyError:
            raise Exception("hydrogen bonded to a selected atom is not part of that selection")
        if side == 1:
            hpos = p1[hl]; dpos = p1[li[c_rep]]; apos = p2[lj[c_rep]]
            donor = g1[c_rep]; acc = g2[c_rep]
        else:
            hpos = p2[hl]; dpos = p2[lj[c_rep]]; apos = p1[li[c_rep]]
            donor = g2[c_rep]; acc = g1[c_rep]
        d_ha = _norm3(hpos - apos)                                   # :166 / :190
        ok = d_ha <= hbondcutoff
        v1 = hpos - apos
        v2 = hpos - dpos
        with np.errstate(invalid="ignore", divide="ignore"):
            ang = np.degrees(np.arccos(np.einsum("ij,ij->i", v1, v2) / (_norm3(v1) * _norm3(v2))))
        ok &= ang >= hbondcutangle                                   # :177 / :201
        for k in np.nonzero(ok)[0]:
            recs.append((int(c_rep[k]), side, int(within[k]), int(donor[k]), int(acc[k]), int(hg[k]),
                         float(d_ha[k]), float(ang[k])))
    recs.sort(key=lambda r: (r[0], r[1], r[2]))
    nc = li.size
    hb_count = np.zeros(nc, np.int64)
    for r in recs:
        hb_count[r[0]] += 1
    hb_ptr = np.zeros(nc + 1, np.int64)
    np.cumsum(hb_count, out=hb_ptr[1:])
    return {
        "n_raw": int(n_raw),
        "idx1": g1, "idx2": g2, "lidx1": li, "lidx2": lj,
        "distance": dist, "weight": weight,
        "hb_ptr": hb_ptr,
        "hb_contact": np.asarray([r[0] for r in recs], np.int64),
        "hb_side": np.asarray([r[1] for r in recs], np.int64),
        "hb_donor": np.asarray([r[3] for r in recs], np.int64),
        "hb_acceptor": np.asarray([r[4] for r in recs], np.int64),
        "hb_hydrogen": np.asarray([r[5] for r in recs], np.int64),
        "hb_distance": np.asarray([r[6] for r in recs], np.float64),
        "hb_angle": np.asarray([r[7] for r in recs], np.float64),
    }


def scan_frames(frames1, frames2, indices1, indices2, cutoff, hbondcutoff, hbondcutangle,
                names, types, bonds, resids=None, segids=None, self_interaction=False, search=None):
    """loop_trajectory_grid over a list of frames (core/multi_trajectory.py:50-222)."""
    hcsr = bonded_hydrogens(names, types, bonds)
    return [frame_contacts(a, b, indices1, indices2, cutoff, hbondcutoff, hbondcutangle, names, hcsr,
                           resids, segids, self_interaction, search)
            for a, b in zip(frames1, frames2)]


# ---------------------------------------------------------------------------------------
# accumulation (core/ContactAnalyzer.py:502-597)
# ---------------------------------------------------------------------------------------
def _key_arrays(amap, g, names, resids, resnames, segids):
    """makeKeyArraysFromMaps (core/ContactAnalyzer.py:107-158) for one atom."""
    out = []
    for k, on in enumerate(amap):
        if on == 1:
            out.append([g, names[g], resids[g], resnames[g], segids[g]][k])
        else:
            out.append("none")
    return out


def _key_string(k1, k2):
    """makeKeyFromKeyArrays (core/ContactAnalyzer.py:251-274)."""
    s = ""
    for tok, item in zip(MAP_TOKENS, k1):
        if item != "none":
            s += tok + str(item)
    s += "-"
    for tok, item in zip(MAP_TOKENS, k2):
        if item != "none":
            s += tok + str(item)
    return s


def accumulate(frames, map1, map2, names, resids, resnames, segids, backbone):
    """analyze_contactResultsWithMaps (core/ContactAnalyzer.py:502-597).

    ``frames`` is the list returned by scan_frames.  Returns
        keys        list[str]  in first-appearance order (:555-556, :582)
        key1, key2  list of key arrays AS STRINGS (the reference round-trips them through
                    makeKeyArraysFromKey, App. A.6 Q7)
        score       float64 (K, F)  scoreArray per key (zero where the key is absent, Q9)
        bb1, sc1, bb2, sc2  float64 (K,)  (:588-591)
        hbonds      int64 (K, F)  hbondFramesScan (core/Biochemistry.py:215-224)
        ncontrib    int64 (K, F)  len(contributingAtoms[f])
    """
    map1 = [int(bool(v)) for v in map1]
    map2 = [int(bool(v)) for v in map2]
    bbset = set(int(b) for b in backbone)
    order, kid = [], {}
    F = len(frames)
    per_frame = []
    for fr in frames:
        acc = {}
        nhb = np.diff(fr["hb_ptr"])
        for c in range(len(fr["idx1"])):
            a, b = int(fr["idx1"][c]), int(fr["idx2"][c])
            k1 = _key_arrays(map1, a, names, resids, resnames, segids)
            k2 = _key_arrays(map2, b, names, resids, resnames, segids)
            key = _key_string(k1, k2)
            if key not in kid:
                kid[key] = len(order)
                order.append((key, k1, k2))
            e = acc.setdefault(key, [0.0, 0.0, 0.0, 0.0, 0.0, 0, 0])
            w = float(fr["weight"][c])
            e[0] += w
            if a in bbset: e[1] += w
            else: e[2] += w
            if b in bbset: e[3] += w
            else: e[4] += w
            e[5] += int(nhb[c])
            e[6] += 1
        per_frame.append(acc)
    K = len(order)
    score = np.zeros((K, F))
    hb = np.zeros((K, F), np.int64)
    ncontrib = np.zeros((K, F), np.int64)
    tot = np.zeros((K, 4))
    for f, acc in enumerate(per_frame):
        for key, e in acc.items():
            k = kid[key]
            score[k, f] = e[0]
            tot[k] += e[1:5]
            hb[k, f] = e[5]
            ncontrib[k, f] = e[6]
    return {
        "keys": [o[0] for o in order],
        "key1": [[str(x) for x in o[1]] for o in order],
        "key2": [[str(x) for x in o[2]] for o in order],
        "score": score,
        "bb1": tot[:, 0], "sc1": tot[:, 1], "bb2": tot[:, 2], "sc2": tot[:, 3],
        "hbonds": hb, "ncontrib": ncontrib,
    }


def hbond_percentage(acc):
    """AccumulatedContact.hbond_percentage per key (core/Biochemistry.py:150-158)."""
    hb = acc["hbonds"]
    return (hb > 0).sum(axis=1) / float(hb.shape[1]) * 100.0


# ---------------------------------------------------------------------------------------
# the same accumulation on arrays (frames with 10^4 .. 10^7 contacts, where the per-contact Python
# loop above takes hours).  tests/test_oracle.py checks it against accumulate() on the fixtures.
# ---------------------------------------------------------------------------------------
def atom_key_codes(amap, n_atoms, names, resids, resnames, segids):
    """int64 code per GLOBAL atom index: equal codes <=> equal key arrays under ``amap``
    (makeKeyArraysFromMaps, core/ContactAnalyzer.py:107-158; the key compares as a string, Q7)."""
    amap = [int(bool(v)) for v in amap]
    cols = []
    if amap[0]:
        cols.append(np.arange(n_atoms, dtype=np.int64))
    for on, vals in ((amap[1], names), (amap[2], resids), (amap[3], resnames), (amap[4], segids)):
        if on:
            _, inv = np.unique(np.asarray([str(v) for v in vals[:n_atoms]], dtype=object).astype(str), return_inverse=True)
            cols.append(inv.astype(np.int64))
    if not cols:
        return np.zeros(n_atoms, np.int64)
    code = np.zeros(n_atoms, np.int64)
    for c in cols:
        _, inv = np.unique(np.stack([code, c], 1), axis=0, return_inverse=True)
        code = inv.reshape(-1).astype(np.int64)
    return code


def accumulate_arrays(frames, code1, code2, backbone, n_atoms):
    """Per frame, the (key -> fscore, bb1score, bb2score, hbonds, contributing contacts) cells of
    core/ContactAnalyzer.py:527-559 as arrays sorted by (code of atom1's key, code of atom2's key)."""
    bb = np.zeros(n_atoms, bool)
    bb[np.asarray(list(backbone), np.int64)] = True
    m2 = int(code2.max()) + 1 if len(code2) else 1
    out = []
    for fr in frames:
        a, b = np.asarray(fr["idx1"], np.int64), np.asarray(fr["idx2"], np.int64)
        k = code1[a] * m2 + code2[b]
        uk, inv = np.unique(k, return_inverse=True)
        w = np.asarray(fr["weight"], np.float64)
        n = uk.size
        out.append({
            "k1": uk // m2, "k2": uk % m2,
            "score": np.bincount(inv, weights=w, minlength=n),
            "bb1": np.bincount(inv, weights=w * bb[a], minlength=n),
            "bb2": np.bincount(inv, weights=w * bb[b], minlength=n),
            "nhb": np.bincount(inv, weights=np.diff(fr["hb_ptr"]), minlength=n).astype(np.int64),
            "ncont": np.bincount(inv, minlength=n).astype(np.int64)})
    return out


def heavy_only_search(is_h1, is_h2, search=None):
    """A ``search`` for frame_contacts that runs the neighbour search on the heavy atoms only.

    Every pair with a hydrogen is dropped right after the search (core/multi_trajectory.py:89), so the
    contact SET is the same; the order inside an atom's row can differ from the reference's (the grid's
    origin is the bounding box of the atoms searched).  For unordered comparisons on frames whose raw
    pair list (10^8 pairs at 3 M atoms) would not fit the host."""
    search = search or (native.ref_find_contacts if native.have_reference() else native.oracle_find_contacts)
    h1 = np.nonzero(~np.asarray(is_h1, bool))[0]
    h2 = np.nonzero(~np.asarray(is_h2, bool))[0]

    def run(xyz1, xyz2, cutoff):
        rp, cols = search(np.ascontiguousarray(xyz1[h1]), np.ascontiguousarray(xyz2[h2]), cutoff)
        row_ptr = np.zeros(xyz1.shape[0] + 1, np.int64)
        row_ptr[h1 + 1] = np.diff(rp)
        np.cumsum(row_ptr, out=row_ptr)
        return row_ptr, h2[cols]
    return run
