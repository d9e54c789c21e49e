"""GPU parity of the SASA and find_within kernels (pc_sasa.cuh) through the C ABI: bit-exact float32 totals,
exposed-point counts and within flags against the compiled reference (oracle/_ref), its golden vectors and
the oracle restatement."""
import os

import numpy as np
import pytest

from oracle import sasa_oracle as so
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
PAIRDIST = 2.0 * (2.0 + 1.4)
RADII = np.array([1.0, 1.5, 1.399999976158142, 1.2999999523162842, 1.899999976158142], np.float32)


def _cloud(n, seed, density=0.1):
    rng = np.random.default_rng(seed)
    L = (max(n, 1) / density) ** (1 / 3)
    pos = (rng.random((n, 3)) * L).astype(np.float32)
    rad = rng.choice(RADII, n).astype(np.float32)
    rl = (rng.random(n) < 0.5).astype(np.int32)
    return pos, rad, rl


def _ubq(rpn11):
    from pycontact_b200.core.Biochemistry import vdwRadius
    ubq = np.nonzero(np.asarray(rpn11.topology.segids) == "UBQ")[0]
    rad = np.array([vdwRadius(rpn11.topology.names[i][0]) for i in ubq], np.float32)
    return ubq, rad


def test_sasa_fixture_equals_reference_golden(rpn11):
    from pycontact_b200.cy_modules.cy_gridsearch import sasa_frames
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_sasa_golden.npz"))
    ubq, rad = _ubq(rpn11)
    coords = np.ascontiguousarray(rpn11.coords[:, ubq])
    free, exposed = sasa_frames(coords, rad, PAIRDIST, 50, 1.4, 1, 0, [0], return_exposed=True)
    assert np.array_equal(free, g["sasa_free"])
    restr = sasa_frames(coords, rad, PAIRDIST, 50, 1.4, 1, 1, g["restricted_list"])
    assert np.array_equal(restr, g["sasa_restricted"])
    spiral = sasa_frames(coords[::10], rad, PAIRDIST, 64, 1.4, 0, 0, [0])
    assert np.array_equal(spiral, g["sasa_spiral64_every10"])
    for f in (0, 31):
        _, _, ex = so.sasa(coords[f], rad, PAIRDIST, 50, 1.4, 1, 0, [0])
        assert np.array_equal(exposed[f], ex)


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (33, 3), (1000, 4), (20000, 5)])
def test_sasa_random_clouds_equal_reference(n, seed):
    from pycontact_b200.cy_modules.cy_gridsearch import sasa_frames
    frames, rads = [], None
    for k in range(3):
        pos, rad, rl = _cloud(n, seed * 10 + k)
        frames.append(pos)
    rads = rad
    coords = np.stack(frames)
    for style, npts in ((1, 50), (0, 37), (1, 200)):
        for restricted in (0, 1):
            got = sasa_frames(coords, rads, PAIRDIST, npts, 1.4, style, restricted, rl)
            for f in range(3):
                assert got[f] == so.ref_sasa(coords[f], rads, PAIRDIST, npts, 1.4, style, restricted, rl)


def test_sasa_dense_sparse_duplicates_and_list_overflow():
    from pycontact_b200.cy_modules.cy_gridsearch import sasa_frames
    # dense blob: > 192 neighbour spheres per atom (several passes over the points per atom)
    pos, rad, rl = _cloud(3000, 21, density=0.6)
    pos[1] = pos[0]
    pos[3] = pos[2] + np.float32(0.01)
    got = sasa_frames(pos[None], rad, 9.0, 50, 1.4, 1, 0, rl)
    assert got[0] == so.ref_sasa(pos, rad, 9.0, 50, 1.4, 1, 0, rl)
    # sparse cloud: far more grid cells than the budget -> the cell edge grows
    pos, rad, rl = _cloud(400, 22, density=1e-6)
    got = sasa_frames(pos[None], rad, PAIRDIST, 50, 1.4, 1, 0, rl)
    assert got[0] == so.ref_sasa(pos, rad, PAIRDIST, 50, 1.4, 1, 0, rl)
    # short pair distance: real neighbours are dropped like the reference drops them
    pos, rad, rl = _cloud(800, 23)
    got = sasa_frames(pos[None], rad, 3.0, 50, 1.4, 1, 1, rl)
    assert got[0] == so.ref_sasa(pos, rad, 3.0, 50, 1.4, 1, 1, rl)
    # no atoms / no frames
    assert sasa_frames(np.zeros((2, 0, 3), np.float32), np.zeros(0, np.float32), PAIRDIST, 50, 1.4, 1, 0, [0]).tolist() == [0, 0]
    assert len(sasa_frames(np.zeros((0, 5, 3), np.float32), np.ones(5, np.float32), PAIRDIST, 50, 1.4, 1, 0, [0])) == 0


def test_cy_sasa_and_worker_signatures(rpn11):
    from pycontact_b200.cy_modules import cy_gridsearch
    from pycontact_b200.sasa import calculate_sasa_parallel, calculate_sasa
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_sasa_golden.npz"))
    ubq, rad = _ubq(rpn11)
    natoms = len(ubq)
    npcoords = np.array(np.reshape(rpn11.coords[3][ubq], (1, natoms * 3)), dtype=np.float32)
    asa = cy_gridsearch.cy_sasa(npcoords, natoms, PAIRDIST, 0, -1, rad, 50, 1.4, 1, 0, np.array([0], np.int32))
    assert isinstance(asa, float) and np.float32(asa) == g["sasa_free"][3]
    with pytest.raises(ValueError):
        cy_gridsearch.cy_sasa(npcoords.astype(np.float64), natoms, PAIRDIST, 0, -1, rad, 50, 1.4, 1, 0, np.array([0], np.int32))
    out = calculate_sasa_parallel([rpn11.coords[f][ubq] for f in range(5)], natoms, PAIRDIST, rad, 50, 1.4, 1, 1,
                                  g["restricted_list"], 0)
    assert [np.float32(v) for v in out] == list(g["sasa_restricted"][:5])
    # the widget's computation on topology + frames, restriction typed like its example
    res = "segid UBQ and same residue as around 5.0 (segid RN11)"
    sasas = calculate_sasa(None, None, "segid UBQ", resseltext=res, nprocs=3, topology=rpn11.topology, frames=rpn11.coords)
    assert np.array_equal(np.array(sasas, np.float32), g["sasa_restricted"])
    area = calculate_sasa(None, None, "segid UBQ", resseltext=res, seltext2="segid UBQ", contact_area=True,
                          topology=rpn11.topology, frames=rpn11.coords[:4])
    assert area == [0.0] * 4


def test_find_within_fixture_and_random(rpn11):
    from pycontact_b200.cy_modules.cy_gridsearch import find_within_frames, cy_find_within
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_sasa_golden.npz"))
    seg = np.asarray(rpn11.topology.segids)
    flg = (seg == "UBQ").astype(np.int32)
    oth = (seg == "RN11").astype(np.int32)
    got = find_within_frames(rpn11.coords, flg, oth, 5.0)
    assert np.array_equal(got, np.unpackbits(g["within5_ubq_of_rn11"], axis=1)[:, :len(seg)])
    assert np.array_equal(got.sum(axis=1), g["within5_counts"])
    rng = np.random.default_rng(7)
    for n, r, dens in ((1, 3.0, 0.1), (50, 2.0, 0.1), (5000, 3.5, 0.1), (60000, 5.0, 0.1), (500, 4.0, 1e-5)):
        pos, _, _ = _cloud(n, n, density=dens)
        f = (rng.random(n) < 0.4).astype(np.int32)
        o = (rng.random(n) < 0.3).astype(np.int32)
        want = so.ref_find_within(pos, f, o, r)
        assert np.array_equal(find_within_frames(pos[None], f, o, r)[0], want)
        z = np.zeros(n, np.int32)
        assert not find_within_frames(pos[None], z, o, r).any()
        assert not find_within_frames(pos[None], f, z, r).any()
        f2 = f.copy()
        res = cy_find_within(pos.reshape(1, -1), f2, o, n, r)
        assert np.array_equal(res, want) and np.array_equal(f2, want)       # flgs overwritten like the reference


def test_around_selection_on_a_frame(rpn11):
    from pycontact_b200.io.selection import select_atoms
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_sasa_golden.npz"))
    want = np.unpackbits(g["within5_ubq_of_rn11"], axis=1)[:, :rpn11.topology.n_atoms]
    for f in (0, 17):
        idx = select_atoms(rpn11.topology, "segid UBQ and around 5.0 (segid RN11)", positions=rpn11.coords[f])
        assert np.array_equal(idx, np.nonzero(want[f])[0])


def test_analyzer_with_around_selection_rescans_every_frame(rpn11):
    """core/ContactAnalyzer.py:319-331: a selection containing `around` is re-evaluated on every frame, so atoms,
    local indices and the search grid change per frame.  Expected values: the oracle run frame by frame on that
    frame's own selection (find_within oracle + residue expansion in numpy)."""
    from oracle import frame_oracle
    from pycontact_b200.core.ContactAnalyzer import Analyzer
    from tests.test_gpu_parity import _compare_traj, RTOL
    topo = rpn11.topology
    F = 12
    coords = rpn11.coords[:F]
    sel2 = "segid UBQ and same residue as around 4.0 (segid RN11)"
    an = Analyzer("x.psf", "x.dcd", 5.0, 2.5, 120, "segid RN11", sel2, topology=topo, frames=coords)
    an.runFrameScan(1)
    assert len(an.contactResults) == F
    seg, res = np.asarray(topo.segids), np.asarray(topo.resids)
    idx1 = np.nonzero(seg == "RN11")[0]
    inner = (seg == "RN11")
    expected, sizes = [], []
    for f in range(F):
        hit = so.find_within(coords[f], (~inner).astype(np.int32), inner.astype(np.int32), 4.0).astype(bool)
        resset = set(zip(seg[hit].tolist(), res[hit].tolist()))
        whole = np.fromiter(((a, b) in resset for a, b in zip(seg.tolist(), res.tolist())), bool, len(seg))
        idx2 = np.nonzero(whole & (seg == "UBQ"))[0]
        sizes.append(len(idx2))
        expected += frame_oracle.scan_frames([coords[f][idx1]], [coords[f][idx2]], idx1, idx2, 5.0, 2.5, 120,
                                             topo.names, topo.types, topo.bonds)
    assert len(set(sizes)) > 1                       # the selection really changes over the frames
    _compare_traj(an.contactResults.traj, expected)
    assert an.scanStats["selection_runs"] > 1
    acc = an.runContactAnalysis([0, 0, 1, 1, 0], [0, 0, 1, 1, 0], 1)
    ora = frame_oracle.accumulate(expected, [0, 0, 1, 1, 0], [0, 0, 1, 1, 0], an.name_array, an.resid_array,
                                  an.resname_array, an.segids, an.backbone)
    assert [a.key1 for a in acc] == ora["key1"] and [a.key2 for a in acc] == ora["key2"]
    np.testing.assert_allclose(np.asarray([a.scoreArray for a in acc], float), ora["score"], rtol=RTOL, atol=1e-12)
    assert np.array_equal(np.asarray([a.hbondFramesScan() for a in acc]), ora["hbonds"])
    # the first frame's selection decides whether the analysis starts at all (:312-313)
    an0 = Analyzer("x.psf", "x.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ and around 0.01 (segid RN11)",
                   topology=topo, frames=coords)
    with pytest.raises(Exception):
        an0.runFrameScan(1)


def test_sasa_and_within_large_frame_equal_reference():
    """300 000 atoms in one frame: many CTAs per frame, cell budget = atoms, one long running sum."""
    from pycontact_b200.cy_modules.cy_gridsearch import sasa_frames, find_within_frames
    pos, rad, rl = _cloud(300000, 77)
    got = sasa_frames(pos[None], rad, PAIRDIST, 50, 1.4, 1, 1, rl)
    assert got[0] == so.ref_sasa(pos, rad, PAIRDIST, 50, 1.4, 1, 1, rl)
    f = rl
    o = (1 - rl).astype(np.int32)
    assert np.array_equal(find_within_frames(pos[None], f, o, 3.0)[0], so.ref_find_within(pos, f, o, 3.0))
