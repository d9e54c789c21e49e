"""CPU tests of the SASA / find_within oracle (oracle/sasa_oracle.py) against the UNMODIFIED reference
compiled into oracle/_ref and against the golden vectors it produced (tests/golden/make_sasa_golden.py),
plus the host-only part of the library (the shared sphere points)."""
import os

import numpy as np
import pytest

from oracle import sasa_oracle as so
from tests.conftest import GOLDEN

needs_ref = pytest.mark.skipif(not so.have_ref(), reason="oracle/_ref is not built")
PAIRDIST = 2.0 * (2.0 + 1.4)
RADII = np.array([1.0, 1.5, 1.399999976158142, 1.2999999523162842, 1.899999976158142], np.float32)


def _cloud(n, seed, density=0.1):
    rng = np.random.default_rng(seed)
    L = (max(n, 1) / density) ** (1 / 3)
    pos = (rng.random((n, 3)) * L).astype(np.float32)
    rad = rng.choice(RADII, n).astype(np.float32)
    rl = (rng.random(n) < 0.5).astype(np.int32)
    return pos, rad, rl


@needs_ref
@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (17, 3), (300, 4), (1200, 5)])
@pytest.mark.parametrize("pointstyle,npts", [(1, 50), (0, 64)])
def test_sasa_oracle_equals_reference(n, seed, pointstyle, npts):
    pos, rad, rl = _cloud(n, seed)
    for restricted in (0, 1):
        tot, areas, exposed = so.sasa(pos, rad, PAIRDIST, npts, 1.4, pointstyle, restricted, rl)
        ref = so.ref_sasa(pos, rad, PAIRDIST, npts, 1.4, pointstyle, restricted, rl)
        assert tot == ref                                       # bit for bit, float32
        assert exposed.max() <= npts


@needs_ref
def test_sasa_oracle_duplicate_atoms_and_short_pairdist():
    # atoms closer than sqrt(0.001) are not neighbours (gridsearch.C:363); a short pairdist drops real ones
    pos, rad, rl = _cloud(200, 11)
    pos[1] = pos[0]
    pos[3] = pos[2] + np.float32(0.01)
    for pd in (PAIRDIST, 3.0):
        tot, _, _ = so.sasa(pos, rad, pd, 50, 1.4, 1, 0, rl)
        assert tot == so.ref_sasa(pos, rad, pd, 50, 1.4, 1, 0, rl)


@needs_ref
@pytest.mark.parametrize("n,seed,r", [(1, 1, 3.0), (50, 2, 2.0), (700, 3, 3.5), (3000, 4, 5.0)])
def test_find_within_oracle_equals_reference(n, seed, r):
    rng = np.random.default_rng(seed)
    pos, _, _ = _cloud(n, seed)
    f = (rng.random(n) < 0.4).astype(np.int32)
    o = (rng.random(n) < 0.3).astype(np.int32)
    assert np.array_equal(so.find_within(pos, f, o, r), so.ref_find_within(pos, f, o, r))
    # empty sets
    z = np.zeros(n, np.int32)
    assert not so.find_within(pos, z, o, r).any() and not so.ref_find_within(pos, z, o, r).any()
    assert not so.find_within(pos, f, z, r).any() and not so.ref_find_within(pos, f, z, r).any()


def test_oracle_reproduces_golden_sasa(rpn11):
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_sasa_golden.npz"))
    ubq = np.nonzero(np.asarray(rpn11.topology.segids) == "UBQ")[0]
    from pycontact_b200.core.Biochemistry import vdwRadius
    rad = np.array([vdwRadius(rpn11.topology.names[i][0]) for i in ubq], np.float32)
    for f in (0, 49):
        tot, _, _ = so.sasa(rpn11.coords[f][ubq], rad, PAIRDIST, 50, 1.4, 1, 0, [0])
        assert tot == g["sasa_free"][f]
        tot, _, _ = so.sasa(rpn11.coords[f][ubq], rad, PAIRDIST, 50, 1.4, 1, 1, g["restricted_list"])
        assert tot == g["sasa_restricted"][f]
    tot, _, _ = so.sasa(rpn11.coords[10][ubq], rad, PAIRDIST, 64, 1.4, 0, 0, [0])
    assert tot == g["sasa_spiral64_every10"][1]


def test_oracle_reproduces_golden_within(rpn11):
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_sasa_golden.npz"))
    seg = np.asarray(rpn11.topology.segids)
    flg = (seg == "UBQ").astype(np.int32)
    oth = (seg == "RN11").astype(np.int32)
    want = np.unpackbits(g["within5_ubq_of_rn11"], axis=1)[:, :len(seg)]
    for f in (0, 25):
        assert np.array_equal(so.find_within(rpn11.coords[f], flg, oth, 5.0), want[f])


def test_library_sphere_points_equal_libc_random_and_libm():
    # host-only entry point: the library's own glibc TYPE_3 generator against libc's srandom / random
    from pycontact_b200 import _lib
    lib = _lib.load()
    for style in (0, 1):
        for npts in (1, 2, 50, 333, 1024):
            out = np.zeros((npts, 3), np.float32)
            assert lib.pc_sasa_points(npts, style, _lib.ptr(out)) == 0
            assert np.array_equal(out, so.sphere_points(npts, style), equal_nan=True)
    assert lib.pc_sasa_points(0, 1, _lib.ptr(np.zeros(3, np.float32))) == _lib.PC_ERR_INVALID


def test_selection_same_residue_and_around_need_positions(rpn11):
    from pycontact_b200.io.selection import select_atoms
    topo = rpn11.topology
    a = select_atoms(topo, "same residue as (segid UBQ and name CA and resid 5:7)")
    seg, res = np.asarray(topo.segids), np.asarray(topo.resids)
    assert np.array_equal(a, np.nonzero((seg == "UBQ") & (res >= 5) & (res <= 7))[0])
    with pytest.raises(NotImplementedError):
        select_atoms(topo, "segid UBQ and around 5.0 (segid RN11)")


def test_select_atoms_frames_host_logic_with_the_oracle_standing_in_for_the_kernel(rpn11, monkeypatch):
    """Parser logic of per-frame selections (boolean operators over (frames x atoms) masks, nested `around`,
    `same residue as` of a per-frame selection) with the find_within ORACLE in place of the CUDA call -- the kernel
    itself is compared with the reference in tests/test_gpu_sasa.py."""
    from pycontact_b200.cy_modules import cy_gridsearch
    from pycontact_b200.io.selection import select_atoms, select_atoms_frames

    def fake(coords, flgs, others, r, device=0):
        return np.stack([so.find_within(c, flgs, others, r) for c in coords])
    monkeypatch.setattr(cy_gridsearch, "find_within_frames", fake)
    topo = rpn11.topology
    frames = rpn11.coords[:3]
    seg, res = np.asarray(topo.segids), np.asarray(topo.resids)
    rn11 = seg == "RN11"
    per = select_atoms_frames(topo, "segid UBQ and around 4.0 (segid RN11)", frames)
    assert len(per) == 3
    for f in range(3):
        want = so.find_within(frames[f], (~rn11).astype(np.int32), rn11.astype(np.int32), 4.0).astype(bool) & (seg == "UBQ")
        assert np.array_equal(per[f], np.nonzero(want)[0])
        assert np.array_equal(select_atoms(topo, "segid UBQ and around 4.0 (segid RN11)", positions=frames[f]), per[f])
    # a static selection comes back once per frame
    stat = select_atoms_frames(topo, "segid UBQ and name CA", frames)
    assert len(stat) == 3 and all(np.array_equal(s, stat[0]) for s in stat)
    # whole residues of a per-frame selection, negation, and an `around` of a per-frame selection
    whole = select_atoms_frames(topo, "same residue as (segid UBQ and around 4.0 (segid RN11))", frames)
    nested = select_atoms_frames(topo, "segid RN11 and not around 3.0 (segid UBQ and around 4.0 (segid RN11))", frames)
    for f in range(3):
        keys = set(zip(seg[per[f]].tolist(), res[per[f]].tolist()))
        want = np.fromiter(((a, b) in keys for a, b in zip(seg.tolist(), res.tolist())), bool, len(seg))
        assert np.array_equal(whole[f], np.nonzero(want)[0])
        inner = np.zeros(len(seg), bool)
        inner[per[f]] = True
        near = so.find_within(frames[f], (~inner).astype(np.int32), inner.astype(np.int32), 3.0).astype(bool)
        assert np.array_equal(nested[f], np.nonzero(rn11 & ~near)[0])
    with pytest.raises(ValueError):
        select_atoms(topo, "all", positions=frames)            # one frame only
