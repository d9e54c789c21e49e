"""Filters and sorts over accumulated contacts, same classes and call signatures as the reference's
``core/ContactFilters.py`` (Operator :4-18, FrameFilter :21-36, NameFilter :39-74, RangeFilter :77-139,
OnlyFilter :153-184, TotalTimeFilter :187-200, ScoreFilter :203-230, SortingOrder / Sorting :233-287).

The reference evaluates every criterion by calling a per-object Python loop (``c.total_time(1, 0)``,
``c.mean_score()``, ``c.hbond_percentage()``, ``c.mean_life_time(...)``) for each contact.  Here the
numeric criteria of ALL contacts come from one ``pc_key_stats`` launch (``pycontact_b200.stats``);
the classes only select / order with the resulting arrays.  Selection semantics are the reference's,
including what passes through unfiltered (keys without the mapped property, non-numeric resids).

Differences kept on purpose: ``TotalTimeFilter.filterContacts`` prints with ``unicode(...)`` in the
reference (:199), a NameError on Python 3 -- not reproduced.
"""
from __future__ import annotations

import numpy as np

from .Biochemistry import *          # noqa: F401,F403  (the reference re-exports Biochemistry here)
from .Biochemistry import AccumulationMapIndex, ContactType
from .. import stats as _stats


class Operator(object):
    """Comparison operator of a filter."""
    greater, smaller, equal, nequal = range(4)
    mapping = {u"greater": greater, u"smaller": smaller, u"equal": equal, u"not equal": nequal}

    def compare(self, value1, value2, operator):
        """Scalar or array ``value1``; None for an unknown operator, like the reference."""
        if operator == self.greater:
            return value1 > value2
        if operator == self.smaller:
            return value1 < value2
        if operator == self.equal:
            return value1 == value2
        if operator == self.nequal:
            return value1 != value2
        return None


def _select(contacts, mask):
    if mask is None:
        return []
    return [c for c, keep in zip(contacts, np.asarray(mask, dtype=bool)) if keep]


class FrameFilter(object):
    """Restricts every contact to a frame range (in place, like the reference)."""

    def __init__(self, name):
        self.name = name

    @staticmethod
    def extractFrameRange(contacts, frameRange):
        lower, upper = frameRange[0], frameRange[1]
        for c in contacts:
            newScores = c.scoreArray[lower:upper]
            newAtoms = c.contributingAtoms[lower:upper]
            c.contributingAtoms = newAtoms
            c.scoreArray = newScores
        return contacts


def _wanted(text):
    return None if text.lower() == u"all" else set(text.split(u","))


class NameFilter(object):
    """Keeps contacts whose mapped property (name, resname, segid ...) is listed."""

    def __init__(self, name):
        self.name = name

    @staticmethod
    def filterContactsByName(contacts, nameA, nameB, mapindex):
        wantA, wantB = _wanted(nameA), _wanted(nameB)
        out = []
        for c in contacts:
            try:
                p1, p2 = c.key1[mapindex], c.key2[mapindex]
            except IndexError:
                out.append(c)              # keys without that property pass
                continue
            if (wantA is None or p1 in wantA) and (wantB is None or p2 in wantB):
                out.append(c)
        return out


def _ranges(text):
    out = []
    for part in text.split(u","):
        lim = part.split(u"-")
        lo = int(lim[0])
        out.append((lo, int(lim[1]) if len(lim) > 1 else lo))
    return out


class RangeFilter(object):
    """Keeps contacts whose integer property (resid, index) lies in the given ranges ("3-10,15")."""

    def __init__(self, name):
        self.name = name

    @staticmethod
    def numberInRanges(number, ranges):
        return any(number in r for r in ranges)

    def filterByRange(self, contacts, residRangeA, residRangeB, mapindex):
        ra = None if residRangeA.lower() == u"all" else _ranges(residRangeA)
        rb = None if residRangeB.lower() == u"all" else _ranges(residRangeB)

        def inside(v, rr):
            return rr is None or any(lo <= v <= hi for lo, hi in rr)

        out = []
        for c in contacts:
            try:
                p1, p2 = int(c.key1[mapindex]), int(c.key2[mapindex])
            except ValueError:
                out.append(c)              # "none" and other non-numeric properties pass
                continue
            if inside(p1, ra) and inside(p2, rb):
                out.append(c)
        return out


class BinaryFilter(object):
    """Base of the filters that compare one number per contact with a value."""

    def __init__(self, name, operator, value):
        self.name = name
        self.operator = Operator.mapping[operator]
        self.value = value

    def filterContacts(self, contacts):
        pass


class OnlyFilter(object):
    """Keeps only contacts of one kind: "hbonds", "hydrophobic", "saltbridges", "other"."""

    def __init__(self, name, operator, value):
        self.name = name
        self.operator = operator
        self.value = value

    def filterContacts(self, contacts):
        if self.operator == "hbonds":
            if len(contacts) == 0:
                return []
            st = _stats.contact_statistics(contacts, assign=False)
            return _select(contacts, st["hbond_percentage"] > 0)
        wanted = {"hydrophobic": ContactType.hydrophobic, "saltbridges": ContactType.saltbr,
                  "other": ContactType.other}.get(self.operator)
        if wanted is None:
            return []
        return [c for c in contacts if c.determine_ctype() == wanted]


class TotalTimeFilter(BinaryFilter):
    """Compares the number of frames with a non-zero score (total_time(1, 0)) with the value."""

    def filterContacts(self, contacts):
        if len(contacts) == 0:
            return []
        st = _stats.contact_statistics(contacts, 1, 0, assign=False)
        return _select(contacts, Operator().compare(st["total_time"], self.value, self.operator))


class ScoreFilter(BinaryFilter):
    """Compares the mean / median score or the H-bond percentage with the value."""
    _column = {u"Mean": "mean", u"Median": "median", u"HB %": "hbond_percentage"}

    def __init__(self, name, operator, value, ftype):
        super(ScoreFilter, self).__init__(name, operator, value)
        self.ftype = ftype

    def filterContacts(self, contacts):
        col = self._column.get(self.ftype)
        if col is None or len(contacts) == 0:
            return []
        st = _stats.contact_statistics(contacts, assign=False)
        for c, m, md in zip(contacts, st["mean"], st["median"]):     # the reference's calls leave these behind
            if col == "mean":
                c.meanScore = float(m)
            elif col == "median":
                c.medianScore = float(md)
        return _select(contacts, Operator().compare(st[col], self.value, self.operator))


class SortingOrder(object):
    ascending, descending = range(2)
    mapping = {u"asc.": ascending, u"desc.": descending}


class Sorting(object):
    """Orders contacts by one key; the time-based keys are evaluated for all contacts in one launch."""

    def __init__(self, name, key, descending):
        self.name = name
        self.key = key
        self.descending = descending
        self.threshold = 0
        self.nspf = 0

    def setThresholdAndNsPerFrame(self, threshold, nspf):
        self.threshold = threshold
        self.nspf = nspf

    def sortContacts(self, contacts):
        key, rev = self.key, self.descending
        if key == u"mean":
            return sorted(contacts, key=lambda c: c.meanScore, reverse=rev)
        if key == u"median":
            return sorted(contacts, key=lambda c: c.medianScore, reverse=rev)
        if key == u"bb/sc type":
            return sorted(contacts, key=lambda c: c.backboneSideChainType, reverse=rev)
        if key == u"contact type":
            return sorted(contacts, key=lambda c: c.contactType, reverse=rev)
        if key in (u"resid 1", u"resid 2"):
            which = "key1" if key == u"resid 1" else "key2"
            try:
                if getattr(contacts[0], which)[AccumulationMapIndex.resid] == "none":
                    return contacts
            except IndexError:
                return []
            return sorted(contacts, key=lambda c: int(getattr(c, which)[AccumulationMapIndex.resid]), reverse=rev)
        column = {u"total time": ("total_time", "ttime"), u"mean lifetime": ("mean_life_time", "meanLifeTime"),
                  u"median lifetime": ("median_life_time", "medianLifeTime")}.get(key)
        if column is None:
            return []
        if len(contacts) == 0:
            return []
        _stats.contact_statistics(contacts, self.nspf, self.threshold, assign=True)
        return sorted(contacts, key=lambda c: getattr(c, column[1]), reverse=rev)
