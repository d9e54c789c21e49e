"""Result types of the contact-analysis path, attribute-compatible with the reference's
``PyContact/core/Biochemistry.py``: ``AtomContact`` (:299-312), ``HydrogenBond`` (:315-331),
``TempContactAccumulate`` (:283-296), ``AccumulatedContact`` (:57-279),
``AccumulationMapIndex`` (:334-340), ``HydrogenBondAtoms`` (:32-33) and the small enums the
accumulated contacts refer to.

The objects are plain Python objects (picklable, same attribute names) so that the
reference's consumers -- filters, exporters, session files -- can read them; the
*containers* that hold them (``pycontact_b200.results``) are array-backed and build
objects on access, because the GPU path produces millions of records per second.
The statistics methods are written on numpy arrays rather than Python loops.
"""
from __future__ import annotations

import numpy as np


# CHARMM radii as VMD has them (core/Biochemistry.py:10-24); keyed by the FIRST character of the atom name
# at the only call site (gui/SasaWidgets.py:185, 190), so "Mg" is never hit -- kept for fidelity.
vdwRadii = {"H": 1.0, "C": 1.5, "N": 1.399999976158142, "O": 1.2999999523162842, "F": 1.47, "Mg": 1.73, "P": 1.8,
            "S": 1.899999976158142}


def vdwRadius(atomType):
    """van der Waals radius of an atom type; 1.5 A when unknown."""
    return vdwRadii.get(atomType, 1.5)


class HydrogenBondAtoms:
    atoms = ["O", "N", "S"]


class AccumulationMapIndex:
    """Positions of the atom properties inside a map / key array, and the key tokens."""
    index, name, resid, resname, segid = range(5)
    mapping = ["i.", "nm.", "r.", "rn.", "s."]
    vmdsel = ["index", "name", "resid", "resname", "segname"]


class ContactType:
    saltbr, hydrophobic, hbond, other = range(4)
    shortcut = ["saltbr", "hydrophobic", "hbond", "other"]


class SideChainPolarity:
    nonpolar, positive, negative, polar, other = range(5)
    mapping = {"nonpolar": nonpolar, "positive": positive, "negative": negative, "polar": polar, "other": other}


class BackboneSidechainType:
    contactsBb, contactsSc = range(2)


class BackboneSidechainContactType:
    bb_only, both, sc_only = range(3)


class AtomContact:
    """One heavy-atom pair inside the cutoff in one frame; indices are GLOBAL atom indices."""
    __slots__ = ("frame", "distance", "weight", "idx1", "idx2", "hbondinfo")

    def __init__(self, frame, distance, weight, idx1, idx2, hbondinfo):
        self.frame = int(frame)
        self.distance = float(distance)
        self.weight = float(weight)
        self.idx1 = int(idx1)
        self.idx2 = int(idx2)
        self.hbondinfo = hbondinfo

    def __getstate__(self):
        return {k: getattr(self, k) for k in self.__slots__}

    def __setstate__(self, state):
        for k, v in state.items():
            setattr(self, k, v)

    def toString(self):
        print("frame: %d, dist: %f, weight: %f, idx1: %d, idx2: %d" % (
            self.frame, self.distance, self.weight, self.idx1, self.idx2))


class HydrogenBond:
    """donor-H ... acceptor triple that passed the distance and angle tests."""
    __slots__ = ("donorIndex", "acceptorIndex", "hydrogenIndex", "acceptorHydrogenDistance", "angle",
                 "usedCutoffDist", "usedCutoffAngle")

    def __init__(self, donorIndex, acceptorIndex, hydrogenIndex, acceptorHydrogenDistance, angle, usedCutoffDist,
                 usedCutoffAngle):
        self.donorIndex = donorIndex
        self.acceptorIndex = acceptorIndex
        self.hydrogenIndex = hydrogenIndex
        self.acceptorHydrogenDistance = acceptorHydrogenDistance
        self.angle = angle
        self.usedCutoffDist = usedCutoffDist
        self.usedCutoffAngle = usedCutoffAngle

    def __getstate__(self):
        return {k: getattr(self, k) for k in self.__slots__}

    def __setstate__(self, state):
        for k, v in state.items():
            setattr(self, k, v)

    def toString(self):
        print("donor: %d, acceptor: %d, hydrogen: %d, dist: %f, angle: %f, usedCutoffDist: %f, usedCutoffAngle: %f"
              % (self.donorIndex, self.acceptorIndex, self.hydrogenIndex, self.acceptorHydrogenDistance, self.angle,
                 self.usedCutoffDist, self.usedCutoffAngle))


class TempContactAccumulate(object):
    """Score of one key in one frame."""

    def __init__(self, key1, key2):
        self.fscore = 0
        self.contributingAtomContacts = []
        self.key1 = key1
        self.key2 = key2
        self.bb1score = 0
        self.bb2score = 0
        self.sc1score = 0
        self.sc2score = 0


def _title_half(key):
    get = lambda k: "" if key[k] == "none" else str(key[k])
    parts = ["%s%s" % (get(AccumulationMapIndex.resname), get(AccumulationMapIndex.resid))]
    for k in (AccumulationMapIndex.index, AccumulationMapIndex.name, AccumulationMapIndex.segid):
        if get(k):
            parts.append("%s %s" % (AccumulationMapIndex.mapping[k], get(k)))
    return " , ".join(p for p in parts if p)


class AccumulatedContact(object):
    """All frames of one accumulation key.

    ``scoreArray`` (one score per frame) and ``contributingAtoms`` (per frame, the
    AtomContacts behind the score) may be handed in lazily by the GPU path: they are then
    materialised on first access.  ``hbondFrames`` comes straight from the device-side
    H-bond counts when available, so ``hbond_percentage`` never has to walk objects.
    """

    def __init__(self, key1, key2):
        self._scores = []
        self._contrib = []
        self._score_src = None          # callable -> sequence of floats
        self._contrib_src = None        # callable -> list[list[AtomContact]]
        self._hb_src = None             # callable -> sequence of ints (H-bonds per frame)
        self.key1 = key1
        self.key2 = key2
        self.title = self.human_readable_title()
        self.bb1 = 0
        self.sc1 = 0
        self.bb2 = 0
        self.sc2 = 0
        self.atom1contactsBy = []
        self.atom2contactsBy = []
        self.backboneSideChainType = []
        self.hbondFrames = []
        self.ttime = 0
        self.meanLifeTime = 0
        self.medianLifeTime = 0
        self.meanScore = 0
        self.medianScore = 0
        self.contactType = 0

    # ---- lazily materialised containers --------------------------------------------------------
    @property
    def scoreArray(self):
        if self._score_src is not None:
            self._scores = list(self._score_src())
            self._score_src = None
        return self._scores

    @scoreArray.setter
    def scoreArray(self, value):
        self._scores, self._score_src = value, None

    @property
    def contributingAtoms(self):
        if self._contrib_src is not None:
            self._contrib = self._contrib_src()
            self._contrib_src = None
        return self._contrib

    @contributingAtoms.setter
    def contributingAtoms(self, value):
        self._contrib, self._contrib_src = value, None

    def __deepcopy__(self, memo):
        # runContactAnalysis returns deepcopy(finalAccumulatedContacts) (core/ContactAnalyzer.py:70):
        # copy what has been materialised, keep sharing the (immutable) array sources of the rest
        new = AccumulatedContact.__new__(AccumulatedContact)
        new.__dict__.update(self.__dict__)
        new.key1, new.key2 = list(self.key1), list(self.key2)
        new._scores = list(self._scores)
        new._contrib = [list(fr) for fr in self._contrib]
        new.hbondFrames = list(self.hbondFrames)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["scoreArray"] = self.scoreArray
        d["contributingAtoms"] = self.contributingAtoms
        if self._hb_src is not None:
            d["hbondFrames"] = [int(v) for v in self._hb_src()]
        for k in ("_scores", "_contrib", "_score_src", "_contrib_src", "_hb_src"):
            d.pop(k, None)
        return d

    def __setstate__(self, d):
        d = dict(d)
        self._scores = d.pop("scoreArray", [])
        self._contrib = d.pop("contributingAtoms", [])
        self._score_src = self._contrib_src = self._hb_src = None
        self.__dict__.update(d)

    # ---- reference API ------------------------------------------------------------------------------
    def getScoreArray(self):
        return self.scoreArray

    def human_readable_title(self):
        return " - ".join(_title_half(k) for k in (self.key1, self.key2))

    def addScore(self, newScore):
        self.scoreArray.append(newScore)

    def addContributingAtoms(self, contAtoms):
        self.contributingAtoms.append(contAtoms)

    def determineBackboneSidechainType(self):
        self.atom1contactsBy = (BackboneSidechainType.contactsBb if self.bb1 > self.sc1
                                else BackboneSidechainType.contactsSc)
        self.atom2contactsBy = (BackboneSidechainType.contactsBb if self.bb2 > self.sc2
                                else BackboneSidechainType.contactsSc)
        pair = (self.atom1contactsBy, self.atom2contactsBy)
        if pair == (BackboneSidechainType.contactsBb, BackboneSidechainType.contactsBb):
            self.backboneSideChainType = BackboneSidechainContactType.bb_only
        elif pair == (BackboneSidechainType.contactsSc, BackboneSidechainType.contactsSc):
            self.backboneSideChainType = BackboneSidechainContactType.sc_only
        else:
            self.backboneSideChainType = BackboneSidechainContactType.both
        return self.backboneSideChainType

    def setScores(self):
        self.mean_score()
        self.median_score()

    def hbondFramesScan(self):
        """H-bonds per frame (sum of len(hbondinfo) over the frame's contributing contacts)."""
        if self._hb_src is not None:
            self.hbondFrames = [int(v) for v in self._hb_src()]
        else:
            self.hbondFrames = [sum(len(c.hbondinfo) for c in frame) for frame in self.contributingAtoms]
        return self.hbondFrames

    def hbond_percentage(self):
        hb = np.asarray(self.hbondFramesScan())
        return float(np.count_nonzero(hb > 0)) / float(len(self.scoreArray)) * 100

    def _above(self, threshold):
        return np.asarray(self.scoreArray, dtype=float) > threshold

    def total_time(self, ns_per_frame, threshold):
        self.ttime = ns_per_frame * int(np.count_nonzero(self._above(threshold))) if len(self.scoreArray) else 0
        return self.ttime

    def life_time(self, ns_per_frame, threshold):
        """Lengths (in ns) of the runs of consecutive frames with score > threshold; like the
        reference, a final entry is always appended for the last frame (0 when the contact is
        not active there)."""
        on = self._above(threshold).astype(np.int8)
        if on.size == 0:
            return []
        edges = np.diff(np.concatenate([[0], on, [0]]))
        starts, ends = np.nonzero(edges == 1)[0], np.nonzero(edges == -1)[0]
        times = [ns_per_frame * int(e - s) for s, e in zip(starts, ends)]
        if not on[-1]:
            times.append(0)
        return times

    def mean_life_time(self, ns_per_frame, threshold):
        self.meanLifeTime = np.mean(self.life_time(ns_per_frame, threshold))
        return self.meanLifeTime

    def median_life_time(self, ns_per_frame, threshold):
        self.medianLifeTime = np.median(self.life_time(ns_per_frame, threshold))
        return self.medianLifeTime

    def mean_score(self):
        total = 0
        for s in self.scoreArray:           # sequential float64 sum, like the reference
            total += s
        self.meanScore = total / len(self.scoreArray)
        return self.meanScore

    def median_score(self):
        self.medianScore = np.median(self.scoreArray)
        return self.medianScore

    def setContactType(self):
        self.contactType = self.determine_ctype()

    def contactTypeAsShortcut(self):
        return ContactType.shortcut[self.determine_ctype()]

    def determine_ctype(self):
        """Contact class from side-chain polarity, bb/sc split and H-bond presence.  The residue
        polarity table is the reference's SQLite table db/aa.db (out of scope here, SURVEY.md
        §2 #10); the 23 residues it holds are inlined."""
        pol = _POLARITY
        p1 = pol.get(str(self.key1[AccumulationMapIndex.resname]).lower(), SideChainPolarity.other)
        p2 = pol.get(str(self.key2[AccumulationMapIndex.resname]).lower(), SideChainPolarity.other)
        self.determineBackboneSidechainType()
        has_hb = any(v > 0 for v in self.hbondFramesScan())
        if (self.atom1contactsBy == BackboneSidechainType.contactsSc
                and self.atom2contactsBy == BackboneSidechainType.contactsSc):
            if {p1, p2} == {SideChainPolarity.positive, SideChainPolarity.negative}:
                return ContactType.saltbr
            if p1 == p2 == SideChainPolarity.nonpolar and not has_hb:
                return ContactType.hydrophobic
            return ContactType.hbond if has_hb else ContactType.other
        return ContactType.hbond if has_hb else ContactType.other


_NP, _PO, _NE, _PL = (SideChainPolarity.nonpolar, SideChainPolarity.positive, SideChainPolarity.negative,
                      SideChainPolarity.polar)
_POLARITY = {
    "gly": _NP, "ala": _NP, "leu": _NP, "ile": _NP, "val": _NP, "phe": _NP, "met": _NP, "pro": _NP,
    "asp": _NE, "glu": _NE, "lys": _PO, "arg": _PO, "hsp": _PO, "asn": _PL, "gln": _PL, "his": _PL,
    "hsd": _PL, "hse": _PL, "cys": _PL, "ser": _PL, "thr": _PL, "tyr": _NP, "trp": _NP,
}


def mean_score_of_contactArray(contacts):
    return np.mean(np.concatenate([np.asarray(c.scoreArray, float) for c in contacts]))


def median_score_of_contactArray(contacts):
    return np.median(np.concatenate([np.asarray(c.scoreArray, float) for c in contacts]))
