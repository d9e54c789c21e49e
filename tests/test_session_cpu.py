"""Session files (core/DataHandler.py:4-28) and the scripting front end (core/Scripting.py) -- SURVEY §8 f-2.

CPU-only: no kernels run here.  The golden session of the reference
(PyContact/exampleData/defaultsession_py3, tests/test_basic.py:32-43 pins 148 contacts and
sum(hbond_percentage) == 676.0 for it) is read where /root/reference exists; everything else
uses small hand-built sessions, so the file works on a box without the reference.
"""
import os
import pickle
import subprocess
import sys
import types

import numpy as np
import pytest

from pycontact_b200.core import Biochemistry as bio
from pycontact_b200.core.DataHandler import DataHandler

REF = "/root/reference"
GOLDEN_SESSION = os.path.join(REF, "PyContact", "exampleData", "defaultsession_py3")
needs_reference = pytest.mark.skipif(not os.path.exists(GOLDEN_SESSION), reason="reference tree not present")


def _tiny_analysis():
    hb = bio.HydrogenBond(3, 17, 4, 1.93, 163.2, 2.5, 120)
    frames = [[bio.AtomContact(0, 3.1, 0.98, 3, 17, [hb]), bio.AtomContact(0, 4.4, 0.12, 5, 17, [])],
              [],
              [bio.AtomContact(2, 3.9, 0.62, 3, 20, [])]]
    acc = []
    for k1, k2, scores, who in ((["none", "none", "12", "GLU", "none"], ["none", "none", "71", "LYS", "none"], [1.10, 0.0, 0.0], 0),
                                (["none", "none", "12", "GLU", "none"], ["none", "none", "72", "ARG", "none"], [0.0, 0.0, 0.62], 2)):
        a = bio.AccumulatedContact(k1, k2)
        a.scoreArray = list(scores)
        a.contributingAtoms = [frames[f] if f == who else [] for f in range(3)]
        a.bb1, a.sc1, a.bb2, a.sc2 = 0.5, 0.6, 0.0, 1.1
        acc.append(a)
    ns = types.SimpleNamespace(psf="a.psf", dcd="a.dcd", cutoff=5.0, hbondcutoff=2.5, hbondcutangle=120.0,
                               sel1text="segid A", sel2text="segid B", contactResults=frames,
                               finalAccumulatedContacts=acc, lastMap1=[0, 0, 1, 1, 0], lastMap2=[0, 0, 1, 1, 0])
    ns.getTrajectoryData = lambda: [["GLU", "LYS"], [12, 71], ["OE1", "NZ"], ["A", "B"], [0], "segid A", "segid B"]
    return ns


def _same_session(a, b):
    ca, aa, ta, ma, ra = a
    cb, ab, tb, mb, rb = b
    assert aa == ab and ta == tb and ma == mb
    assert len(ca) == len(cb) and len(ra) == len(rb)
    for x, y in zip(ca, cb):
        assert x.key1 == y.key1 and x.key2 == y.key2 and x.title == y.title
        assert list(x.scoreArray) == list(y.scoreArray)
        assert (x.bb1, x.sc1, x.bb2, x.sc2) == (y.bb1, y.sc1, y.bb2, y.sc2)
        assert x.hbond_percentage() == y.hbond_percentage()
        assert [len(f) for f in x.contributingAtoms] == [len(f) for f in y.contributingAtoms]
    for fa, fb in zip(ra, rb):
        assert [(c.frame, c.idx1, c.idx2, c.distance, c.weight, len(c.hbondinfo)) for c in fa] == \
               [(c.frame, c.idx1, c.idx2, c.distance, c.weight, len(c.hbondinfo)) for c in fb]


@pytest.mark.parametrize("reference_layout", [True, False])
def test_round_trip(tmp_path, reference_layout):
    ana = _tiny_analysis()
    path = str(tmp_path / "s.session")
    DataHandler.writeSessionToFile(path, ana, reference_layout=reference_layout)
    got = DataHandler.importSessionFromFile(path)
    want = [ana.finalAccumulatedContacts,
            [ana.psf, ana.dcd, ana.cutoff, ana.hbondcutoff, ana.hbondcutangle, ana.sel1text, ana.sel2text],
            ana.getTrajectoryData(), [ana.lastMap1, ana.lastMap2], ana.contactResults]
    _same_session(got, want)
    hb = got[4][0][0].hbondinfo[0]
    assert (hb.donorIndex, hb.acceptorIndex, hb.hydrogenIndex) == (3, 17, 4)


def test_reference_layout_names_reference_classes(tmp_path):
    """A session written with reference_layout=True must load in a process that only has
    ``PyContact.core.Biochemistry`` -- emulated with plain stand-in classes (no __setstate__, like the
    reference's), in a child interpreter that cannot import this package."""
    path = str(tmp_path / "ref.session")
    DataHandler.writeSessionToFile(path, _tiny_analysis(), reference_layout=True)
    raw = open(path, "rb").read()
    assert b"pycontact_b200" not in raw and b"PyContact.core.Biochemistry" in raw
    child = r'''
import pickle, sys, types
pkg = types.ModuleType("PyContact"); core = types.ModuleType("PyContact.core"); b = types.ModuleType("PyContact.core.Biochemistry")
for n in ("AtomContact", "HydrogenBond", "AccumulatedContact", "TempContactAccumulate"):
    setattr(b, n, type(n, (object,), {"__module__": "PyContact.core.Biochemistry"}))
sys.modules.update({"PyContact": pkg, "PyContact.core": core, "PyContact.core.Biochemistry": b})
d = pickle.load(open(sys.argv[1], "rb"))
c = d["contacts"][0]
assert type(c).__module__ == "PyContact.core.Biochemistry"
assert c.scoreArray == [1.10, 0.0, 0.0] and c.key2[3] == "LYS" and c.title
a = d["analyzer"][-1][0][0]
assert (a.idx1, a.idx2, a.frame) == (3, 17, 0) and a.hbondinfo[0].hydrogenIndex == 4
assert len(c.contributingAtoms) == 3 and c.contributingAtoms[0][0].idx1 == 3
print("ok")
'''
    out = subprocess.run([sys.executable, "-S", "-c", child, path], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


@needs_reference
def test_reads_the_reference_golden_session():
    contacts, arguments, trajArgs, maps, contactResults = DataHandler.importSessionFromFile(GOLDEN_SESSION)
    assert len(contacts) == 148                                   # tests/test_basic.py:38
    assert sum(c.hbond_percentage() for c in contacts) == 676.0   # tests/test_basic.py:43
    assert len(contactResults) == 50 and sum(len(f) for f in contactResults) == 23208
    assert arguments[2:] == [5.0, 2.5, 120.0, "segid RN11", "segid UBQ"]
    assert all(isinstance(c, bio.AccumulatedContact) for c in contacts)
    assert isinstance(contactResults[0][0], bio.AtomContact)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rpn11_ubq_golden.npz"))
    assert [c.title for c in contacts] == list(g["acc_title"])
    assert np.array_equal(np.array([c.scoreArray for c in contacts]), g["acc_score"])
    assert np.array_equal(np.array([c.hbond_percentage() for c in contacts]), g["acc_hbond_percentage"])


@needs_reference
def test_reference_reads_our_session(tmp_path):
    """Golden session -> our objects -> our file -> the UNMODIFIED reference classes."""
    contacts, arguments, trajArgs, maps, contactResults = DataHandler.importSessionFromFile(GOLDEN_SESSION)
    ana = types.SimpleNamespace(psf=arguments[0], dcd=arguments[1], cutoff=arguments[2], hbondcutoff=arguments[3],
                                hbondcutangle=arguments[4], sel1text=arguments[5], sel2text=arguments[6],
                                contactResults=contactResults, finalAccumulatedContacts=contacts,
                                lastMap1=maps[0], lastMap2=maps[1], getTrajectoryData=lambda: trajArgs)
    path = str(tmp_path / "ours.session")
    DataHandler.writeSessionToFile(path, ana)
    child = r'''
import pickle, sys
sys.path.insert(0, "/root/reference")
from PyContact.core.DataHandler import DataHandler
import PyContact.core.Biochemistry as B
contacts, arguments, trajArgs, maps, contactResults = DataHandler.importSessionFromFile(sys.argv[1])
assert len(contacts) == 148 and isinstance(contacts[0], B.AccumulatedContact)
assert sum(c.hbond_percentage() for c in contacts) == 676.0
assert sum(len(f) for f in contactResults) == 23208 and isinstance(contactResults[0][0], B.AtomContact)
print("ok")
'''
    out = subprocess.run([sys.executable, "-c", child, path], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_scripting_surface():
    from pycontact_b200.core import Scripting
    cfg = Scripting.JobConfig(5.0, 2.5, 120, [0, 0, 1, 1, 0], [0, 0, 1, 1, 0], "segid RN11", "segid UBQ")
    job = Scripting.PyContactJob("a.psf", "a.dcd", "title", cfg)
    assert (job.topo, job.traj, job.name, job.analyzer) == ("a.psf", "a.dcd", "title", None)
    assert (cfg.cutoff, cfg.hbondcutoff, cfg.hbondcutangle, cfg.sel1, cfg.sel2) == (5.0, 2.5, 120, "segid RN11", "segid UBQ")
    assert hasattr(Scripting, "Analyzer") and hasattr(Scripting, "DataHandler")
