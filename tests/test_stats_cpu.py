"""Per-contact statistics and filters, CPU part: the oracle against golden vectors made with the
reference's own classes, and the filters that need no statistics against the reference's filter classes."""
import os
import sys

import numpy as np
import pytest

from oracle import stats_oracle
from pycontact_b200.core import Biochemistry as bio
from pycontact_b200.core import ContactFilters as cf

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "rpn11_ubq_stats_golden.npz")
REF = "/root/reference"


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_stats_oracle_matches_reference_classes(golden, tag):
    ns, thr = golden["params_" + tag]
    st = stats_oracle.key_statistics(golden["score"], golden["hbonds"], ns if tag == "b" else 1, thr if tag == "b" else 0)
    # same operations in the same order as core/Biochemistry.py:150-213 -> bit-identical
    assert np.array_equal(st["mean"], golden["mean"])
    assert np.array_equal(st["median"], golden["median"])
    assert np.array_equal(st["hbond_percentage"], golden["hbond_percentage"])
    assert np.array_equal(st["total_time"], golden["total_time_" + tag])
    assert np.array_equal(st["mean_life_time"], golden["mean_life_time_" + tag])
    assert np.array_equal(st["median_life_time"], golden["median_life_time_" + tag])
    assert np.array_equal(st["life_entries"], golden["life_entries_" + tag])
    assert golden["hbond_percentage"].sum() == 676.0          # tests/test_basic.py:43


def test_life_time_quirks():
    lt = stats_oracle.life_time
    assert lt([0, 0, 0], 1, 0) == [0]                       # never active: the trailing entry only
    assert lt([1, 1, 1], 1, 0) == [3]                       # active to the end: the open run is the trailing entry
    assert lt([1, 0, 1, 1, 0], 1, 0) == [1, 2, 0]           # a drop on the last frame appends the run AND a zero
    assert lt([0, 1, 0, 1], 2, 0) == [2, 2]
    assert lt([], 1, 0) == []


def _contacts(n=12, seed=3):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        c = bio.AccumulatedContact(["none", "none", str(10 + k % 5), ["GLU", "LYS", "ALA"][k % 3], ["A", "B"][k % 2]],
                                   ["none", "none", str(70 - k), ["ARG", "ASP"][k % 2], "B"])
        c.scoreArray = list(rng.random(6) * (rng.random(6) > 0.4))
        c.contributingAtoms = [[] for _ in range(6)]
        c.meanScore = float(np.mean(c.scoreArray))
        c.medianScore = float(np.median(c.scoreArray))
        out.append(c)
    return out


def _ref_filters():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import PyContact.core.ContactFilters as rcf
    return rcf


def _titles(cs):
    return [(c.key1, c.key2) for c in cs]


def test_name_range_frame_filters_match_reference():
    rcf = _ref_filters()
    idx = bio.AccumulationMapIndex
    for a, b, mi in (("GLU", "all", idx.resname), ("GLU,ALA", "ARG", idx.resname), ("all", "all", idx.resname),
                     ("all", "B", idx.segid), ("A", "all", idx.segid), ("XXX", "all", idx.resname)):
        assert _titles(cf.NameFilter.filterContactsByName(_contacts(), a, b, mi)) == \
               _titles(rcf.NameFilter.filterContactsByName(_contacts(), a, b, mi))
    for a, b in (("10-12", "all"), ("all", "60-65,69"), ("11", "59-70"), ("all", "all")):
        assert _titles(cf.RangeFilter("r").filterByRange(_contacts(), a, b, idx.resid)) == \
               _titles(rcf.RangeFilter("r").filterByRange(_contacts(), a, b, idx.resid))
    # a non-numeric property passes
    assert len(cf.RangeFilter("r").filterByRange(_contacts(), "1-2", "all", idx.resname)) == 12
    ours = cf.FrameFilter.extractFrameRange(_contacts(), [1, 4])
    theirs = rcf.FrameFilter.extractFrameRange(_contacts(), [1, 4])
    assert [c.scoreArray for c in ours] == [c.scoreArray for c in theirs]
    assert all(len(c.contributingAtoms) == 3 for c in ours)


def test_sorting_without_statistics_matches_reference():
    rcf = _ref_filters()
    for key in (u"mean", u"median", u"resid 1", u"resid 2"):
        for desc in (False, True):
            assert _titles(cf.Sorting("s", key, desc).sortContacts(_contacts())) == \
                   _titles(rcf.Sorting("s", key, desc).sortContacts(_contacts()))
    assert cf.Sorting("s", u"resid 1", False).sortContacts([]) == []
    assert cf.Sorting("s", u"no such key", False).sortContacts(_contacts()) == []


def test_operator():
    op = cf.Operator()
    assert cf.Operator.mapping == {u"greater": 0, u"smaller": 1, u"equal": 2, u"not equal": 3}
    assert op.compare(2, 1, op.greater) and op.compare(1, 2, op.smaller) and op.compare(1, 1, op.equal)
    assert op.compare(1, 2, op.nequal) and op.compare(1, 2, 17) is None
    assert op.compare(np.array([1.0, 3.0]), 2.0, op.greater).tolist() == [False, True]
