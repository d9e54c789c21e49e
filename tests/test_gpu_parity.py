"""CUDA path vs the oracle / golden vectors -- needs a B200 (``-m gpu``).

Everything goes through the C ABI (ctypes -> libpycontact_b200.so).  The bar
(BASELINE.json north_star): contact pair sets and H-bond triples bit-exact, including
the reference's in-frame order; distances, weights, angles and scores within 1e-5
relative (the device stores float32 records, the reference float64).
"""
import os

import numpy as np
import pytest

from oracle import frame_oracle, native
from pycontact_b200 import prepare, synth
from pycontact_b200.core.ContactAnalyzer import Analyzer
from pycontact_b200.core.multi_trajectory import loop_trajectory_grid, run_load_parallel
from pycontact_b200.cy_modules.cy_gridsearch import cy_find_contacts, find_contacts_csr
from pycontact_b200.engine import ContactEngine
from pycontact_b200.results import TrajectoryContacts
from pycontact_b200.scheduler import ArraySource, run_scan

pytestmark = pytest.mark.gpu
RTOL = 1e-5          # north_star: "scores and angles agree within 1e-5 relative tolerance (FP32 vs FP64)"


def _selections(topo, idx1, idx2, backbone):
    hcsr = prepare.hydrogen_csr_from_bond_table(topo.names, topo.types, topo.bonds)
    bb = np.zeros(topo.n_atoms, bool)
    bb[backbone] = True
    seg = prepare.string_codes(topo.segids)
    s1 = prepare.prepare_selection(idx1, topo.names, topo.resids, seg, bb, hcsr)
    s2 = None if idx2 is None else prepare.prepare_selection(idx2, topo.names, topo.resids, seg, bb, hcsr)
    return s1, s2


def _compare_traj(traj: TrajectoryContacts, oracle_frames):
    assert traj.n_frames == len(oracle_frames)
    for f, fr in enumerate(oracle_frames):
        lo, hi = traj.frame_ptr[f], traj.frame_ptr[f + 1]
        assert np.array_equal(traj.idx1[lo:hi], fr["idx1"]), "frame %d: idx1 / order differs" % f
        assert np.array_equal(traj.idx2[lo:hi], fr["idx2"]), "frame %d: idx2 / order differs" % f
        np.testing.assert_allclose(traj.distance[lo:hi], fr["distance"], rtol=RTOL)
        np.testing.assert_allclose(traj.weight[lo:hi], fr["weight"], rtol=RTOL, atol=1e-30)
        hlo, hhi = traj.hb_ptr[lo], traj.hb_ptr[hi]
        assert np.array_equal(traj.hb_ptr[lo:hi + 1] - hlo, fr["hb_ptr"]), "frame %d: H-bond flags differ" % f
        assert np.array_equal(traj.hb_donor[hlo:hhi], fr["hb_donor"])
        assert np.array_equal(traj.hb_acceptor[hlo:hhi], fr["hb_acceptor"])
        assert np.array_equal(traj.hb_hydrogen[hlo:hhi], fr["hb_hydrogen"])
        np.testing.assert_allclose(traj.hb_distance[hlo:hhi], fr["hb_distance"], rtol=RTOL)
        np.testing.assert_allclose(traj.hb_angle[hlo:hhi], fr["hb_angle"], rtol=RTOL)


def _golden_frames(g, prefix=""):
    fp, hp = g[prefix + "frame_ptr"], g[prefix + "hb_ptr"]
    out = []
    for k in range(len(fp) - 1):
        lo, hi = fp[k], fp[k + 1]
        hlo, hhi = hp[lo], hp[hi]
        out.append({"idx1": g[prefix + "idx1"][lo:hi], "idx2": g[prefix + "idx2"][lo:hi],
                    "distance": g[prefix + "distance"][lo:hi], "weight": g[prefix + "weight"][lo:hi],
                    "hb_ptr": hp[lo:hi + 1] - hlo, "hb_donor": g[prefix + "hb_donor"][hlo:hhi],
                    "hb_acceptor": g[prefix + "hb_acceptor"][hlo:hhi], "hb_hydrogen": g[prefix + "hb_hydrogen"][hlo:hhi],
                    "hb_distance": g[prefix + "hb_distance"][hlo:hhi], "hb_angle": g[prefix + "hb_angle"][hlo:hhi]})
    return out


def _check_accumulated(acc_objs, g, prefix="acc_"):
    assert len(acc_objs) == len(g[prefix + "key1"])
    assert [a.key1 for a in acc_objs] == g[prefix + "key1"].tolist()         # first-appearance ORDER
    assert [a.key2 for a in acc_objs] == g[prefix + "key2"].tolist()
    assert [a.title for a in acc_objs] == g[prefix + "title"].tolist()
    score = np.asarray([a.scoreArray for a in acc_objs], float)
    np.testing.assert_allclose(score, g[prefix + "score"], rtol=RTOL, atol=1e-12)
    assert np.array_equal(score == 0, g[prefix + "score"] == 0)
    for k in ("bb1", "sc1", "bb2", "sc2"):
        np.testing.assert_allclose([getattr(a, k) for a in acc_objs], g[prefix + k], rtol=RTOL, atol=1e-9)
    assert np.array_equal(np.asarray([a.hbondFramesScan() for a in acc_objs]), g[prefix + "hbonds"])
    assert np.array_equal(np.asarray([[len(fr) for fr in a.contributingAtoms] for a in acc_objs]), g[prefix + "ncontrib"])
    np.testing.assert_allclose([a.hbond_percentage() for a in acc_objs], g[prefix + "hbond_percentage"])


# ---------------------------------------------------------------------------------------------------
# B1: cy_find_contacts
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n1,n2,cutoff,seed", [(500, 700, 5.0, 1), (2000, 300, 4.6, 2), (1500, 1500, 7.3, 3),
                                               (50, 4000, 3.0, 4), (1, 1, 5.0, 5), (300, 300, 12.0, 6),
                                               (20000, 30000, 5.0, 7)])
def test_cy_find_contacts_matches_oracle_rows_and_order(n1, n2, cutoff, seed):
    a, b, _, _ = synth.cloud(n1, n2, seed=seed)
    rp, cols = find_contacts_csr(a.reshape(1, -1), n1, b.reshape(1, -1), n2, cutoff)
    orp, ocols = native.oracle_find_contacts(a, b, cutoff)
    assert np.array_equal(rp, orp)
    assert np.array_equal(cols, ocols)


def test_cy_find_contacts_return_shape_and_fixture(rpn11, golden_ab):
    s1, s2 = golden_ab["sel1"], golden_ab["sel2"]
    for f in (0, 31):
        x1 = np.ascontiguousarray(rpn11.coords[f, s1].reshape(1, -1))
        x2 = np.ascontiguousarray(rpn11.coords[f, s2].reshape(1, -1))
        rows = cy_find_contacts(x1, len(s1), x2, len(s2), 5.0)
        assert len(rows) == len(s1) + len(s2) and all(len(r) == 0 for r in rows[len(s1):])   # Q3
        orp, ocols = native.oracle_find_contacts(x1, x2, 5.0)
        assert rows[:len(s1)] == native.as_list_of_lists(orp, ocols, 0)
        assert sum(len(r) for r in rows) == golden_ab["n_raw"][f]
    with pytest.raises(ValueError):
        cy_find_contacts(x1.astype(np.float64), len(s1), x2, len(s2), 5.0)


def test_cy_find_contacts_self_keeps_ii_pairs():
    a, _, _, _ = synth.cloud(400, 1, seed=9)
    rows = cy_find_contacts(a.reshape(1, -1), 400, a.copy().reshape(1, -1), 400, 5.0)
    assert all(i in rows[i] for i in range(400))


# ---------------------------------------------------------------------------------------------------
# B2 / B3: frame scan + accumulation on the reference's fixture (golden session)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path,ctas", [(1, 0), (1, 1), (1, 2), (1, 3), (1, 4), (2, 0)])
def test_scan_fixture_matches_golden_session(rpn11, golden_ab, path, ctas):
    idx1, idx2 = golden_ab["sel1"].astype(np.int64), golden_ab["sel2"].astype(np.int64)
    s1, s2 = _selections(rpn11.topology, idx1, idx2, rpn11.backbone)
    with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
        eng.set_path(path)
        for sl in eng.slots:
            eng.lib.pc_set_small_ctas(sl.ctx, ctas)
        src = ArraySource(np.ascontiguousarray(rpn11.coords[:, idx1]), np.ascontiguousarray(rpn11.coords[:, idx2]))
        traj, _, stats = run_scan([eng], src, batch_frames=16, order_keys=True)
    assert stats["path"] == path
    assert stats["contacts"] == 23208 and stats["hbonds"] == 457
    assert traj.packed_sorts > 0                     # the device's order keys took the one-pass host sort (results.py)
    _compare_traj(traj, _golden_frames(golden_ab))


def test_analyzer_single_core_like_reference_test(rpn11, golden_ab):
    """tests/test_basic.py:32-43 of the reference, plus element-wise parity with its golden session."""
    an = Analyzer("rpn11_ubq.psf", "rpn11_ubq.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ",
                  topology=rpn11.topology, frames=rpn11.coords)
    an.runFrameScan(1)
    assert len(an.contactResults) == 50
    map1 = [0, 0, 1, 1, 0]
    map2 = [0, 0, 1, 1, 0]
    an.runContactAnalysis(map1, map2, 1)
    assert len(an.finalAccumulatedContacts) == 148
    hbond_sum = 0
    for c in an.finalAccumulatedContacts:
        hbond_sum += c.hbond_percentage()
    assert hbond_sum == 676.0
    _check_accumulated(an.finalAccumulatedContacts, golden_ab)
    # object view of a frame
    fr = an.contactResults[7]
    lo = golden_ab["frame_ptr"][7]
    assert [c.idx1 for c in fr] == golden_ab["idx1"][lo:lo + len(fr)].tolist()
    assert all(c.frame == 7 for c in fr)
    assert an.backbone == rpn11.backbone.tolist()
    # a second accumulation with other maps reuses the scan (name-level keys)
    acc2 = an.runContactAnalysis([0, 1, 1, 1, 1], [0, 1, 1, 1, 1], 1)
    ora = frame_oracle.accumulate(_as_oracle_frames(golden_ab), [0, 1, 1, 1, 1],
                                  [0, 1, 1, 1, 1], an.name_array, an.resid_array, an.resname_array, an.segids,
                                  an.backbone)
    assert [a.key1 for a in acc2] == ora["key1"] and [a.key2 for a in acc2] == ora["key2"]
    np.testing.assert_allclose(np.asarray([a.scoreArray for a in acc2], float), ora["score"], rtol=RTOL, atol=1e-12)


def _as_oracle_frames(g, prefix=""):
    out = _golden_frames(g, prefix)
    for fr in out:
        fr["idx1"] = fr["idx1"].astype(np.int64)
        fr["idx2"] = fr["idx2"].astype(np.int64)
    return out


def test_analyzer_self_interaction(rpn11, golden_self):
    """tests/test_basic.py:45-51 (runs, 50 frames) + the oracle-derived values for this mode."""
    an = Analyzer("rpn11_ubq.psf", "rpn11_ubq.dcd", 5.0, 2.5, 120, "segid RN11", "self",
                  topology=rpn11.topology, frames=rpn11.coords)
    an.runFrameScan(1)
    assert len(an.contactResults) == 50
    traj = an.contactResults.traj
    assert np.diff(traj.frame_ptr).tolist() == golden_self["frame_counts"].tolist()
    assert np.diff(traj.hb_ptr[traj.frame_ptr]).tolist() == golden_self["frame_hbonds"].tolist()
    assert (traj.idx1 != traj.idx2).all()
    wsum = np.add.reduceat(traj.weight, traj.frame_ptr[:-1])
    np.testing.assert_allclose(wsum, golden_self["frame_weight_sum"], rtol=RTOL)
    frames = golden_self["full_frames"].tolist()
    gold = _golden_frames(golden_self, "full_")
    for k, f in enumerate(frames):
        lo, hi = traj.frame_ptr[f], traj.frame_ptr[f + 1]
        assert np.array_equal(traj.idx1[lo:hi], gold[k]["idx1"]) and np.array_equal(traj.idx2[lo:hi], gold[k]["idx2"])
    acc = an.runContactAnalysis([0, 0, 1, 1, 0], [0, 0, 1, 1, 0], 1)
    assert len(acc) == 788
    assert sum(a.hbond_percentage() for a in acc) == 5212.0
    _check_accumulated(acc, golden_self)


def test_zero_atomselection_raises(rpn11):
    """tests/test_basic.py:53-63."""
    an = Analyzer("x.psf", "x.dcd", 5.0, 2.5, 120, "segid A", "resid 10000",
                  topology=rpn11.topology, frames=rpn11.coords)
    with pytest.raises(Exception):
        an.runFrameScan(1)
    with pytest.raises(Exception):
        an.runFrameScan(4)


def test_loop_trajectory_grid_signature_and_q1(rpn11, golden_ab):
    """B2: the chunk-level entry with the reference's argument shapes."""
    t = rpn11.topology
    idx1, idx2 = golden_ab["sel1"].astype(np.int64), golden_ab["sel2"].astype(np.int64)
    ptr, bid = t.bonds_of_atoms()

    class Bonds:
        def __init__(self, a):
            ids = bid[ptr[a]:ptr[a + 1]]
            self.types = [(t.types[i], t.types[j]) for i, j in t.bonds[ids]]
            self._idx = [np.asarray(t.bonds[b]) for b in ids]

        def to_indices(self):
            return self._idx
    bonds = [Bonds(a) for a in range(t.n_atoms)]
    frames = list(range(0, 50, 7))
    res = loop_trajectory_grid([rpn11.coords[f, idx1] for f in frames], [rpn11.coords[f, idx2] for f in frames],
                               [idx1] * len(frames), [idx2] * len(frames), [5.0, 2.5, 120],
                               [bonds, t.names], False)
    assert len(res) == len(frames)
    gold = _golden_frames(golden_ab)
    for k, f in enumerate(frames):
        assert [c.idx1 for c in res[k]] == gold[f]["idx1"].tolist()
        assert [c.idx2 for c in res[k]] == gold[f]["idx2"].tolist()
        assert all(c.frame == 0 for c in res[k])                          # Q1
        assert [len(c.hbondinfo) for c in res[k]] == np.diff(gold[f]["hb_ptr"]).tolist()


# ---------------------------------------------------------------------------------------------------
# synthetic systems vs the oracle on the same seeded inputs, both kernel paths
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,path,self_mode,nframes", [
    ("tiny", 1, False, 6), ("tiny", 2, False, 6), ("tiny", 1, True, 6), ("tiny", 2, True, 6),
    ("small", 1, False, 8), ("small", 2, False, 8), ("small", 2, True, 4),
    ("c3_small", 0, False, 4), ("c3_small", 2, False, 4), ("c4_small", 0, True, 2), ("c4_small", 2, True, 2)])
def test_synthetic_matches_oracle(name, path, self_mode, nframes):
    s = synth.make_system(name)
    t = s.topology
    frames = list(range(nframes))
    coords = synth.frames_host(s, frames)
    idx1 = s.sel1
    idx2 = s.sel1 if self_mode else s.sel2
    ora = frame_oracle.scan_frames([coords[k, idx1] for k in range(nframes)], [coords[k, idx2] for k in range(nframes)],
                                   idx1, idx2, 5.0, 2.5, 120, t.names, t.types, t.bonds, t.resids, t.segids, self_mode)
    s1, s2 = _selections(t, idx1, None if self_mode else idx2, s.backbone)
    with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
        eng.set_path(path)
        src = ArraySource(np.ascontiguousarray(coords[:, idx1]), None if self_mode else np.ascontiguousarray(coords[:, idx2]))
        gm1 = prepare.group_map(idx1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
        gm2 = prepare.group_map(idx2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
        eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
        traj, table, stats = run_scan([eng], src, batch_frames=3, order_keys=True, accumulate=True,
                                      group_maps=(gm1, gm2))
    _compare_traj(traj, ora)
    acc = frame_oracle.accumulate(ora, s.maps[0], s.maps[1], t.names, t.resids, t.resnames, t.segids, s.backbone)
    keys = [table.key_arrays(k, t.names, t.resids, t.resnames, t.segids) for k in range(table.n_keys)]
    assert [k[0] for k in keys] == acc["key1"] and [k[1] for k in keys] == acc["key2"]
    np.testing.assert_allclose(table.score_matrix(), acc["score"], rtol=RTOL, atol=1e-12)
    assert np.array_equal(table.hbond_matrix(), acc["hbonds"])
    np.testing.assert_allclose(table.bb1, acc["bb1"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(table.sc2, acc["sc2"], rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("self_mode", [False, True])
def test_small_path_one_pass_prediction(self_mode):
    """A CTA of the small-frame kernel keeps atoms for its NEXT frame's grid during the bounding-box pass (box of its
    previous frame + a margin) and skips the second pass over the frame when the exact box lies inside it.  One launch
    over more frames than there are CTAs (2 x 148), so that CTAs see a second and a third frame: contained boxes
    (thermal jitter), boxes that jump (whole system translated: second pass), empty grids in between (selection 2 moved
    away) -- every frame equals the oracle, with and without the fused accumulation."""
    s = synth.make_system("small")
    t = s.topology
    nframes = 700
    frames = list(range(nframes))
    coords = synth.frames_host(s, frames).copy()
    rng = np.random.default_rng(5)
    jump = rng.random(nframes) < 0.3
    coords[jump] += rng.uniform(-40, 40, (int(jump.sum()), 1, 3)).astype(np.float32)      # the box is somewhere else
    idx1 = s.sel1
    idx2 = s.sel1 if self_mode else s.sel2
    if not self_mode:
        far = np.nonzero(rng.random(nframes) < 0.05)[0]
        coords[np.ix_(far, idx2)] += np.float32(500.0)                                    # no overlap: empty grid
    ora = frame_oracle.scan_frames([coords[k, idx1] for k in range(nframes)], [coords[k, idx2] for k in range(nframes)],
                                   idx1, idx2, 5.0, 2.5, 120, t.names, t.types, t.bonds, t.resids, t.segids, self_mode)
    s1, s2 = _selections(t, idx1, None if self_mode else idx2, s.backbone)
    gm1 = prepare.group_map(idx1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
    gm2 = prepare.group_map(idx2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
    with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
        eng.set_path(1)
        eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
        src = ArraySource(np.ascontiguousarray(coords[:, idx1]), None if self_mode else np.ascontiguousarray(coords[:, idx2]))
        traj, table, _ = run_scan([eng], src, batch_frames=nframes, order_keys=True, accumulate=True, group_maps=(gm1, gm2))
        _compare_traj(traj, ora)
        _, table2, _ = run_scan([eng], src, batch_frames=nframes, order_keys=False, accumulate=True, group_maps=(gm1, gm2))
    acc = frame_oracle.accumulate(ora, s.maps[0], s.maps[1], t.names, t.resids, t.resnames, t.segids, s.backbone)
    np.testing.assert_allclose(table.score_matrix(), acc["score"], rtol=RTOL, atol=1e-12)
    # fused accumulation (no order keys): same keys in another order
    o1, o2 = np.argsort(table.keys), np.argsort(table2.keys)
    assert np.array_equal(table.keys[o1], table2.keys[o2])
    np.testing.assert_allclose(table2.score_matrix()[o2], table.score_matrix()[o1], rtol=RTOL, atol=1e-9)
    assert np.array_equal(table2.hbond_matrix()[o2], table.hbond_matrix()[o1])


@pytest.mark.parametrize("cutoff", [3.0, 4.6, 8.0, 12.0])
def test_cutoff_sweep_small_system(cutoff):
    """C5-style sweep on a small solvated system: pair sets exact at every cutoff."""
    s = synth.make_system("small")
    t = s.topology
    coords = synth.frames_host(s, [0, 1])
    ora = frame_oracle.scan_frames([coords[k, s.sel1] for k in range(2)], [coords[k, s.sel2] for k in range(2)],
                                   s.sel1, s.sel2, cutoff, 2.5, 120, t.names, t.types, t.bonds)
    s1, s2 = _selections(t, s.sel1, s.sel2, s.backbone)
    with ContactEngine(s1, s2, cutoff, 2.5, 120) as eng:
        src = ArraySource(np.ascontiguousarray(coords[:, s.sel1]), np.ascontiguousarray(coords[:, s.sel2]))
        traj, _, _ = run_scan([eng], src, order_keys=True)
    _compare_traj(traj, ora)


def test_missing_hydrogen_raises_like_reference(rpn11, golden_ab):
    """A hydrogen bonded to a selected N/O/S atom that is not in the selection: the reference
    fails at core/multi_trajectory.py:158; the device flags it."""
    from pycontact_b200._lib import PcError
    idx1, idx2 = golden_ab["sel1"].astype(np.int64), golden_ab["sel2"].astype(np.int64)
    names = rpn11.topology.names
    heavy_only = np.asarray([i for i in idx1 if not names[i].startswith("H")], np.int64)
    s1, s2 = _selections(rpn11.topology, heavy_only, idx2, rpn11.backbone)
    with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
        with pytest.raises(PcError):
            eng.scan(np.ascontiguousarray(rpn11.coords[:4, heavy_only]), np.ascontiguousarray(rpn11.coords[:4, idx2]))


def test_device_generator_frames_are_analysed_exactly():
    """bench.py analyses frames generated ON the device; parity is checked on exactly those bytes."""
    from pycontact_b200 import device_synth
    s = synth.make_system("small")
    t = s.topology
    gen = device_synth.DeviceTrajectory(s, device=0)
    host = gen.frames_to_host([0, 5, 63])
    assert host.shape == (3, t.n_atoms, 3) and np.isfinite(host).all()
    np.testing.assert_allclose(host, synth.frames_host(s, [0, 5, 63]), atol=2.5)       # same system, other RNG
    ora = frame_oracle.scan_frames([host[k, s.sel1] for k in range(3)], [host[k, s.sel2] for k in range(3)],
                                   s.sel1, s.sel2, 5.0, 2.5, 120, t.names, t.types, t.bonds)
    s1, s2 = _selections(t, s.sel1, s.sel2, s.backbone)
    with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
        src = ArraySource(np.ascontiguousarray(host[:, s.sel1]), np.ascontiguousarray(host[:, s.sel2]))
        traj, _, _ = run_scan([eng], src, order_keys=True)
    _compare_traj(traj, ora)


@pytest.mark.parametrize("name,self_mode", [("small", False), ("small", True), ("tiny", False)])
def test_fused_accumulation_equals_standalone_kernel(name, self_mode):
    """The scan kernel's in-kernel accumulation (no order keys) and pc_accumulate give the same rows."""
    s = synth.make_system(name)
    t = s.topology
    coords = synth.frames_host(s, range(5))
    idx1 = s.sel1
    idx2 = s.sel1 if self_mode else s.sel2
    s1, s2 = _selections(t, idx1, None if self_mode else idx2, s.backbone)
    gm1 = prepare.group_map(idx1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
    gm2 = prepare.group_map(idx2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
    tables = []
    for fused, ctas in ((1, 2), (0, 2), (1, 1)):
        with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
            eng.set_path(1)
            for sl in eng.slots:
                eng.lib.pc_set_fused_accumulate(sl.ctx, fused)
                eng.lib.pc_set_small_ctas(sl.ctx, ctas)
            eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
            src = ArraySource(np.ascontiguousarray(coords[:, idx1]), None if self_mode else np.ascontiguousarray(coords[:, idx2]))
            _, table, _ = run_scan([eng], src, batch_frames=2, order_keys=False, accumulate=True, group_maps=(gm1, gm2))
        tables.append(table)
    a, b, c1 = tables
    assert np.array_equal(np.sort(a.keys), np.sort(c1.keys))
    # the fused table sums in float32 (order-dependent in the last bits), the stand-alone kernel in float64
    np.testing.assert_allclose(a.score_matrix()[np.argsort(a.keys)], c1.score_matrix()[np.argsort(c1.keys)], rtol=2e-6)
    assert a.n_keys == b.n_keys and a.n_keys > 0
    oa, ob = np.argsort(a.keys), np.argsort(b.keys)
    assert np.array_equal(a.keys[oa], b.keys[ob])
    np.testing.assert_allclose(a.score_matrix()[oa], b.score_matrix()[ob], rtol=2e-6, atol=1e-300)
    assert np.array_equal(a.hbond_matrix()[oa], b.hbond_matrix()[ob])
    np.testing.assert_allclose(a.bb1[oa], b.bb1[ob], rtol=2e-6, atol=1e-300)
    np.testing.assert_allclose(a.bb2[oa], b.bb2[ob], rtol=2e-6, atol=1e-300)


def test_packed_rows_equal_full_rows():
    """pc_fetch_packed_rows (the 24-byte transfer format) carries the same cells as pc_fetch."""
    s = synth.make_system("small")
    t = s.topology
    coords = synth.frames_host(s, range(6))
    s1, s2 = _selections(t, s.sel1, s.sel2, s.backbone)
    gm1 = prepare.group_map(s.sel1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
    gm2 = prepare.group_map(s.sel2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
    from pycontact_b200._lib import DeviceBuffer
    from pycontact_b200.results import AccumulationTable
    with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
        eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
        d1 = DeviceBuffer.from_array(np.ascontiguousarray(coords[:, s.sel1]))
        d2 = DeviceBuffer.from_array(np.ascontiguousarray(coords[:, s.sel2]))
        eng.scan_device(d1.ptr, d2.ptr, 6, accumulate=True)
        full, rs, rc = (a.copy() for a in eng.fetch_rows(0))
        packed, ps, pc = (a.copy() for a in eng.fetch_rows(0, packed=True))
        eng.set_pack_format(12)
        slim, ss, sc = (a.copy() for a in eng.fetch_rows(0, packed=True))
    assert slim.dtype.itemsize == 12 and np.array_equal(slim["key"], full["key"].astype(np.uint32))
    assert np.array_equal(slim["score"], packed["score"]) and np.array_equal(slim["nhbonds"], packed["nhbonds"])
    assert full.size == packed.size > 0 and np.array_equal(rs[:6], ps[:6]) and np.array_equal(rc[:6], pc[:6])
    assert np.array_equal(full["key"], packed["key"])
    np.testing.assert_allclose(packed["score"], full["score"], rtol=2e-7)
    np.testing.assert_allclose(packed["bb1"], full["bb1"], rtol=2e-7)
    assert np.array_equal(packed["nhbonds"], full["nhbonds"]) and np.array_equal(packed["ncontacts"], full["ncontacts"])
    ta, tb = AccumulationTable(6, gm1, gm2), AccumulationTable(6, gm1, gm2)
    ta.add_packed(packed, ps[:6], pc[:6], 0)
    frame_of = np.zeros(full.size, np.int64)
    for f in range(6):
        frame_of[rs[f]:rs[f] + rc[f]] = f
    tb.add_rows(full, frame_of)
    ta.finalize(); tb.finalize()
    oa, ob = np.argsort(ta.keys), np.argsort(tb.keys)
    np.testing.assert_allclose(ta.score_matrix()[oa], tb.score_matrix()[ob], rtol=2e-7)


def test_run_load_parallel_returns_reference_shape(rpn11, golden_ab):
    """run_load_parallel (core/multi_trajectory.py:354-455): [allContacts, resname, resid, name, segids, backbone],
    contacts labelled frame 0 like the reference's pool path (Q1)."""
    res = run_load_parallel(1, "rpn11_ubq.psf", "rpn11_ubq.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ",
                            topology=rpn11.topology, frames=rpn11.coords[:6])
    contacts, resname, resid, name, segids, backbone = res
    assert len(contacts) == 6 and len(resname) == len(resid) == len(name) == len(segids) == 4327
    assert backbone == rpn11.backbone.tolist()
    fp = golden_ab["frame_ptr"]
    assert [len(fr) for fr in contacts] == np.diff(fp[:7]).tolist()
    assert all(c.frame == 0 for c in contacts[3])
    assert [c.idx2 for c in contacts[5]] == golden_ab["idx2"][fp[5]:fp[6]].tolist()


def test_analyzer_two_gpus_equals_one(rpn11, golden_ab):
    """The analogue of test_multiCore_analysis (tests/test_basic.py:73-84): frames sharded over 2 GPUs."""
    from pycontact_b200 import _lib
    if _lib.load().pc_device_count() < 2:
        pytest.skip("needs two GPUs")
    an = Analyzer("rpn11_ubq.psf", "rpn11_ubq.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ",
                  topology=rpn11.topology, frames=rpn11.coords)
    an.runFrameScan(2)
    assert len(an.contactResults) == 50
    an.runContactAnalysis([0, 0, 1, 1, 0], [0, 0, 1, 1, 0], 2)
    assert len(an.finalAccumulatedContacts) == 148
    assert sum(c.hbond_percentage() for c in an.finalAccumulatedContacts) == 676.0
    _check_accumulated(an.finalAccumulatedContacts, golden_ab)


def test_analyzer_from_psf_dcd_files(tmp_path):
    """The file-based constructor path (PSF + DCD readers, selection strings) end to end."""
    from tests.test_io_cpu import _write_dcd, _write_psf
    s = synth.make_system("tiny")
    t = s.topology
    coords = synth.frames_host(s, range(6))
    psf, dcd = str(tmp_path / "tiny.psf"), str(tmp_path / "tiny.dcd")
    _write_psf(psf, t)
    _write_dcd(dcd, coords)
    an = Analyzer(psf, dcd, 5.0, 2.5, 120, s.sel1text, s.sel2text, batch_frames=4)
    seen = []
    an.frameUpdate.connect(seen.append)
    an.runFrameScan(1)
    acc = an.runContactAnalysis(s.maps[0], s.maps[1], 1)
    ora = frame_oracle.scan_frames([coords[f, s.sel1] for f in range(6)], [coords[f, s.sel2] for f in range(6)],
                                   s.sel1, s.sel2, 5.0, 2.5, 120, t.names, t.types, t.bonds)
    _compare_traj(an.contactResults.traj, ora)
    oacc = frame_oracle.accumulate(ora, s.maps[0], s.maps[1], t.names, t.resids, t.resnames, t.segids, s.backbone)
    assert [a.key1 for a in acc] == oacc["key1"] and [a.key2 for a in acc] == oacc["key2"]
    np.testing.assert_allclose(np.asarray([a.scoreArray for a in acc], float), oacc["score"], rtol=RTOL, atol=1e-12)
    assert seen and seen[-1] == 1.0
    assert an.getFilePaths() == (psf, dcd) and an.lastMap1 == s.maps[0]


def test_analyzer_from_psf_xtc_files(tmp_path):
    """The same constructor path on an XTC trajectory (the reference's tests open one, tests/test_basic.py:28-30): the
    file's quantised coordinates (nm * 1000, rounded; back in Angstrom) are what the oracle gets."""
    from tests.test_io_cpu import _write_psf, _write_xtc
    from pycontact_b200.io.xtc import XTCReader
    s = synth.make_system("tiny")
    t = s.topology
    psf, xtc = str(tmp_path / "tiny.psf"), str(tmp_path / "tiny.xtc")
    _write_psf(psf, t)
    _write_xtc(xtc, synth.frames_host(s, range(5)) / np.float32(10.0))
    rd = XTCReader(xtc)
    coords = np.stack([rd.frame(f) for f in range(len(rd))])
    an = Analyzer(psf, xtc, 5.0, 2.5, 120, s.sel1text, s.sel2text, batch_frames=2)
    an.runFrameScan(1)
    acc = an.runContactAnalysis(s.maps[0], s.maps[1], 1)
    ora = frame_oracle.scan_frames([coords[f, s.sel1] for f in range(5)], [coords[f, s.sel2] for f in range(5)],
                                   s.sel1, s.sel2, 5.0, 2.5, 120, t.names, t.types, t.bonds)
    _compare_traj(an.contactResults.traj, ora)
    oacc = frame_oracle.accumulate(ora, s.maps[0], s.maps[1], t.names, t.resids, t.resnames, t.segids, s.backbone)
    assert [a.key1 for a in acc] == oacc["key1"] and [a.key2 for a in acc] == oacc["key2"]
    np.testing.assert_allclose(np.asarray([a.scoreArray for a in acc], float), oacc["score"], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("path", [1, 2])
def test_edge_cases_empty_results_and_zero_distance(path):
    """Selections of hydrogens only, far-apart selections (empty grid), single atoms, coincident atoms."""
    from pycontact_b200.io.psf import Topology
    names = ["N", "HN", "O", "H1", "CA", "OH2"]
    n = len(names)
    topo = Topology(["A"] * 3 + ["B"] * 3, np.arange(1, n + 1), ["ALA"] * n, names, ["NH1", "H", "O", "HT", "CT1", "OT"],
                    np.zeros(n), np.ones(n), np.asarray([[0, 1], [5, 3]], np.int64))
    xyz = np.asarray([[0, 0, 0], [1, 0, 0], [3, 0, 0], [50, 0, 0], [0, 0, 0], [2.9, 0, 0]], np.float32)
    frames = np.stack([xyz, xyz + np.float32(0.25)])

    def scan(i1, i2):
        s1, s2 = _selections(topo, np.asarray(i1, np.int64), np.asarray(i2, np.int64), [])
        with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
            eng.set_path(path)
            src = ArraySource(np.ascontiguousarray(frames[:, i1]), np.ascontiguousarray(frames[:, i2]))
            traj, _, st = run_scan([eng], src, order_keys=True)
        ora = frame_oracle.scan_frames([frames[f, i1] for f in range(2)], [frames[f, i2] for f in range(2)],
                                       np.asarray(i1), np.asarray(i2), 5.0, 2.5, 120, topo.names, topo.types, topo.bonds)
        _compare_traj(traj, ora)
        return traj
    assert scan([1], [3]).idx1.size == 0                       # hydrogens only: every pair is dropped
    assert scan([0, 1], [3]).idx1.size == 0                    # heavy atom vs a far hydrogen
    t = scan([0, 1, 2], [4, 5, 3])                             # coincident atoms N / CA: distance 0 is a contact
    assert (t.distance == 0).sum() == 2 and t.idx1.size == 8
    assert t.hb_donor.size > 0                                 # OH2-H1 ... wait: N-HN...OH2 geometry gives H-bonds
    far = frames.copy()
    far[:, 3:] += np.float32(500.0)
    src_far = (np.ascontiguousarray(far[:, [0, 1, 2]]), np.ascontiguousarray(far[:, [4, 5, 3]]))
    s1, s2 = _selections(topo, np.asarray([0, 1, 2]), np.asarray([4, 5, 3]), [])
    with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
        eng.set_path(path)
        traj, _, st = run_scan([eng], ArraySource(*src_far), order_keys=True)
    assert traj.idx1.size == 0 and st["contacts"] == 0        # bounding boxes do not touch: empty grid


def test_sparse_fold_equals_dense_fold_and_host_sums():
    """pc_fold_rows_sparse (hash table across batches) == pc_fold_rows (dense table) == sums over the rows."""
    import ctypes
    from pycontact_b200 import _lib
    from pycontact_b200._lib import DeviceBuffer, check
    s = synth.make_system("small")
    t = s.topology
    coords = synth.frames_host(s, range(8))
    s1, s2 = _selections(t, s.sel1, s.sel2, s.backbone)
    gm1 = prepare.group_map(s.sel1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
    gm2 = prepare.group_map(s.sel2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
    n_keys = gm1.n_groups * gm2.n_groups
    lib = _lib.load()
    with ContactEngine(s1, s2, 5.0, 2.5, 120, slots=2) as eng:
        eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
        eng.fold_sparse_init(1 << 14)
        dense = DeviceBuffer.from_array(np.zeros(n_keys * 4))
        host = np.zeros((n_keys, 4))
        for b, sid in ((0, 0), (1, 1)):
            d1 = DeviceBuffer.from_array(np.ascontiguousarray(coords[4 * b:4 * b + 4][:, s.sel1]))
            d2 = DeviceBuffer.from_array(np.ascontiguousarray(coords[4 * b:4 * b + 4][:, s.sel2]))
            eng.submit_device(sid, d1.ptr, d2.ptr, 4, accumulate=True, fold="sparse")
            eng.collect_device(sid)
            check(lib.pc_fold_rows(eng.slots[sid].ctx, dense.ptr, n_keys, eng.slots[sid].stream))
            rows, _, _ = eng.fetch_rows(sid)
            k = rows["key"].astype(np.int64)
            np.add.at(host[:, 0], k, rows["score"]); np.add.at(host[:, 1], k, rows["bb1"])
            np.add.at(host[:, 2], k, rows["bb2"]); np.add.at(host[:, 3], k, (rows["nhbonds"] > 0) * 1.0)
        ent = eng.fold_sparse_fetch()
        dn = dense.download(np.float64, n_keys * 4).reshape(n_keys, 4)
    np.testing.assert_allclose(dn, host, rtol=1e-12, atol=1e-300)      # same rows folded two ways
    assert ent.size == np.count_nonzero(host[:, 0]) and np.all(np.diff(ent["key"].astype(np.int64)) > 0)
    k = ent["key"].astype(np.int64)
    np.testing.assert_allclose(np.stack([ent["score"], ent["bb1"], ent["bb2"], ent["hbframes"]], 1), host[k], rtol=1e-12)


def test_session_file_from_a_gpu_analysis(rpn11, golden_ab, tmp_path):
    """core/DataHandler.py:20-28 on a live analysis: the lazy, array-backed containers of the GPU path are
    expanded into the reference's session layout; re-imported, the session reproduces the pinned numbers of
    tests/test_basic.py:38-43 and the golden per-contact data."""
    from pycontact_b200.core.DataHandler import DataHandler
    an = Analyzer("rpn11_ubq.psf", "rpn11_ubq.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ",
                  topology=rpn11.topology, frames=rpn11.coords)
    an.runFrameScan(1)
    an.runContactAnalysis([0, 0, 1, 1, 0], [0, 0, 1, 1, 0], 1)
    path = str(tmp_path / "gpu.session")
    DataHandler.writeSessionToFile(path, an)
    assert b"pycontact_b200" not in open(path, "rb").read()
    contacts, arguments, trajArgs, maps, contactResults = DataHandler.importSessionFromFile(path)
    assert len(contacts) == 148 and sum(c.hbond_percentage() for c in contacts) == 676.0
    assert arguments == ["rpn11_ubq.psf", "rpn11_ubq.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ"]
    assert maps == [[0, 0, 1, 1, 0], [0, 0, 1, 1, 0]] and len(trajArgs) == 7
    assert [len(f) for f in contactResults] == np.diff(golden_ab["frame_ptr"]).tolist()
    flat = [c for f in contactResults for c in f]
    assert [c.idx1 for c in flat] == golden_ab["idx1"].tolist() and [c.idx2 for c in flat] == golden_ab["idx2"].tolist()
    np.testing.assert_allclose([c.distance for c in flat], golden_ab["distance"], rtol=RTOL)
    np.testing.assert_allclose([c.weight for c in flat], golden_ab["weight"], rtol=RTOL)
    assert sum(len(c.hbondinfo) for c in flat) == 457
    _check_accumulated(contacts, golden_ab)


def test_scripting_job_on_files(tmp_path):
    """core/Scripting.py:28-41: PyContactJob.runJob + writeSessionToFile on PSF/DCD files."""
    from pycontact_b200.core.DataHandler import DataHandler
    from pycontact_b200.core.Scripting import JobConfig, PyContactJob
    from tests.test_io_cpu import _write_dcd, _write_psf
    s = synth.make_system("tiny")
    coords = synth.frames_host(s, range(5))
    psf, dcd = str(tmp_path / "tiny.psf"), str(tmp_path / "tiny.dcd")
    _write_psf(psf, s.topology)
    _write_dcd(dcd, coords)
    job = PyContactJob(psf, dcd, str(tmp_path / "tinyjob"), JobConfig(5.0, 2.5, 120, s.maps[0], s.maps[1], s.sel1text, s.sel2text))
    job.runJob(1)
    job.writeSessionToFile()
    contacts, arguments, trajArgs, maps, contactResults = DataHandler.importSessionFromFile(str(tmp_path / "tinyjob.session"))
    t = s.topology
    ora = frame_oracle.scan_frames([coords[f, s.sel1] for f in range(5)], [coords[f, s.sel2] for f in range(5)],
                                   s.sel1, s.sel2, 5.0, 2.5, 120, t.names, t.types, t.bonds)
    oacc = frame_oracle.accumulate(ora, s.maps[0], s.maps[1], t.names, t.resids, t.resnames, t.segids, s.backbone)
    assert [a.key1 for a in contacts] == oacc["key1"] and [a.key2 for a in contacts] == oacc["key2"]
    np.testing.assert_allclose(np.asarray([a.scoreArray for a in contacts], float), oacc["score"], rtol=RTOL, atol=1e-12)
    assert [len(f) for f in contactResults] == [len(f["idx1"]) for f in ora]
    assert arguments[:2] == [psf, dcd] and maps == [s.maps[0], s.maps[1]]


# ---- per-contact statistics and filters on the device ---------------------------------
STATS_RTOL = 1e-12      # float64 sums in a different order / count * ns instead of repeated addition


def _check_stats(st, ora, exact_counts=True):
    for k in ("mean", "median", "hbond_percentage", "total_time", "mean_life_time", "median_life_time"):
        np.testing.assert_allclose(st[k], ora[k], rtol=STATS_RTOL, atol=1e-300, err_msg=k)
    assert np.array_equal(st["life_entries"], ora["life_entries"])
    assert np.array_equal(st["frames_above"], ora["frames_above"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_key_stats_against_reference_golden(tag):
    """pc_key_stats vs the values the reference's AccumulatedContact methods give on its golden session."""
    from pycontact_b200 import stats
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rpn11_ubq_stats_golden.npz"))
    ns, thr = g["params_" + tag]
    st = stats.key_statistics(g["score"], g["hbonds"], ns, thr)
    np.testing.assert_allclose(st["mean"], g["mean"], rtol=STATS_RTOL)
    assert np.array_equal(st["median"], g["median"])
    assert np.array_equal(st["hbond_percentage"], g["hbond_percentage"]) and st["hbond_percentage"].sum() == 676.0
    np.testing.assert_allclose(st["total_time"], g["total_time_" + tag], rtol=STATS_RTOL)
    np.testing.assert_allclose(st["mean_life_time"], g["mean_life_time_" + tag], rtol=STATS_RTOL)
    np.testing.assert_allclose(st["median_life_time"], g["median_life_time_" + tag], rtol=STATS_RTOL)
    assert np.array_equal(st["life_entries"], g["life_entries_" + tag])


def _random_scores(K, F, seed):
    rng = np.random.default_rng(seed)
    sc = rng.random((K, F))
    # on/off episodes of random length so that runs of every size occur
    gate = np.repeat(rng.random((K, F // 7 + 1)) > 0.45, 7, axis=1)[:, :F]
    sc = sc * gate
    sc[0] = 0.0
    if K > 1:
        sc[1] = 0.9
    hb = (rng.random((K, F)) > 0.8).astype(np.int64) * rng.integers(1, 4, (K, F)) * (sc > 0)
    return sc, hb


@pytest.mark.parametrize("F", [1, 2, 3, 50, 777, 4096])
def test_key_stats_against_oracle(F):
    from oracle import stats_oracle
    from pycontact_b200 import stats
    sc, hb = _random_scores(37, F, 100 + F)
    for ns, thr in ((1, 0), (0.02, 0.5)):
        _check_stats(stats.key_statistics(sc, hb, ns, thr), stats_oracle.key_statistics(sc, hb, ns, thr))


def test_key_stats_full_size_and_limits():
    """C2's 10^4 frames (and the 16384-frame limit): oracle on a sample of keys, array identities on all."""
    from oracle import stats_oracle
    from pycontact_b200 import stats
    for K, F in ((600, 10000), (9, 16384)):
        sc, hb = _random_scores(K, F, 7)
        st = stats.key_statistics(sc, hb, 0.02, 0.5)
        sample = np.unique(np.r_[0, 1, np.random.default_rng(1).integers(0, K, 6)])
        ora = stats_oracle.key_statistics(sc[sample], hb[sample], 0.02, 0.5)
        _check_stats({k: v[sample] for k, v in st.items()}, ora)
        above = sc > 0.5
        assert np.array_equal(st["frames_above"], above.sum(1))
        np.testing.assert_allclose(st["mean"], sc.mean(1), rtol=1e-12)
        assert np.array_equal(st["median"], np.median(sc, axis=1))
        assert np.array_equal(st["hbond_percentage"], (hb > 0).sum(1) / F * 100)
        starts = above[:, :1].sum(1) + (above[:, 1:] & ~above[:, :-1]).sum(1)          # maximal runs
        assert np.array_equal(st["life_entries"], starts + (~above[:, -1]).astype(int))
        np.testing.assert_allclose(st["mean_life_time"] * st["life_entries"], st["total_time"], rtol=1e-12)
    with pytest.raises(ValueError):
        stats.key_statistics(np.zeros((2, 16385)), None)
    empty = stats.key_statistics(np.zeros((0, 10)), None)
    assert all(v.shape == (0,) for v in empty.values())
    nohb = stats.key_statistics(np.ones((3, 5)), None, 1, 0)
    assert nohb["hbond_percentage"].tolist() == [0, 0, 0] and nohb["median_life_time"].tolist() == [5, 5, 5]


def _cells_of(sc, hb, rng):
    """(keys x frames) matrices -> CSR over keys of their non-zero cells (plus a few explicit zero cells), shuffled inside a key."""
    kp, fr, val, hbv = [0], [], [], []
    for k in range(sc.shape[0]):
        f = np.nonzero((sc[k] != 0) | (hb[k] > 0) | (rng.random(sc.shape[1]) < 0.02))[0]
        f = rng.permutation(f)
        fr.append(f); val.append(sc[k, f]); hbv.append(hb[k, f])
        kp.append(kp[-1] + f.size)
    return np.asarray(kp), np.concatenate(fr), np.concatenate(val), np.concatenate(hbv)


@pytest.mark.parametrize("F", [1, 2, 50, 777, 10000, 16384])
def test_key_stats_from_cells_equals_dense_and_oracle(F):
    """pc_key_stats_rows assembles every key's row in shared memory from the (frame, key) cells: same table as the dense
    call (bit for bit: the medians are order statistics, the sums run over the same values) and as the oracle."""
    from oracle import stats_oracle
    from pycontact_b200 import stats
    rng = np.random.default_rng(F)
    K = 23
    sc, hb = _random_scores(K, F, 300 + F)
    if K > 3 and F > 2:
        sc[2] = 0.7; sc[2, F // 2] = 0.0                  # two long runs
        sc[3] = rng.normal(0, 1, F)                       # negative scores: the select's key order
    kp, fr, val, hbv = _cells_of(sc, hb, rng)
    for ns, thr in ((1, 0), (0.02, 0.5)):
        rows = stats.key_statistics_rows(kp, fr, val, hbv, F, ns, thr)
        dense = stats.key_statistics(sc, hb, ns, thr)
        for k in rows:
            if k == "mean":
                np.testing.assert_allclose(rows[k], dense[k], rtol=STATS_RTOL, atol=1e-300)
            else:
                assert np.array_equal(rows[k], dense[k]), k
        sample = np.arange(min(K, 6))
        _check_stats({k: v[sample] for k, v in rows.items()}, stats_oracle.key_statistics(sc[sample], hb[sample], ns, thr))
    assert np.array_equal(dense["median"], np.median(sc, axis=1))
    empty = stats.key_statistics_rows(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0), None, 10)
    assert all(v.shape == (0,) for v in empty.values())
    none = stats.key_statistics_rows(np.zeros(3, np.int64), np.zeros(0, np.int32), np.zeros(0), None, 10, 1, 0)   # keys without cells
    assert none["mean"].tolist() == [0, 0] and none["life_entries"].tolist() == [1, 1]


def test_accumulation_table_statistics_on_the_fixture(rpn11, golden_ab):
    """AccumulationTable.statistics (cells -> pc_key_stats_rows) on the reference's fixture vs the values the reference's
    AccumulatedContact methods give on its golden session."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rpn11_ubq_stats_golden.npz"))
    a = Analyzer("rpn11_ubq.psf", "rpn11_ubq.dcd", 5.0, 2.5, 120, "segid RN11", "segid UBQ",
                 topology=rpn11.topology, frames=rpn11.coords)
    a.runFrameScan(1)
    a.runContactAnalysis([0, 0, 1, 1, 0], [0, 0, 1, 1, 0], 1)
    ns, thr = g["params_a"]
    st = a._table.statistics(ns, thr)
    np.testing.assert_allclose(st["mean"], g["mean"], rtol=1e-6)
    np.testing.assert_allclose(st["median"], g["median"], rtol=1e-6, atol=1e-12)
    assert np.array_equal(st["hbond_percentage"], g["hbond_percentage"]) and st["hbond_percentage"].sum() == 676.0
    np.testing.assert_allclose(st["total_time"], g["total_time_a"], rtol=1e-9)
    assert np.array_equal(st["life_entries"], g["life_entries_a"])


def test_filters_and_sorting_on_device():
    """core/ContactFilters.py:187-287: the criteria come from pc_key_stats; expectations from the oracle."""
    from oracle import stats_oracle
    from pycontact_b200.core import Biochemistry as bio
    from pycontact_b200.core import ContactFilters as cf
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rpn11_ubq_stats_golden.npz"))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "rpn11_ubq_golden.npz"))

    def contacts():
        out = []
        for k in range(len(g["score"])):
            c = bio.AccumulatedContact(list(gold["acc_key1"][k]), list(gold["acc_key2"][k]))
            c.scoreArray = g["score"][k].tolist()
            c.contributingAtoms = [[] for _ in range(g["score"].shape[1])]
            c.hbondFrames = g["hbonds"][k].tolist()
            c.hbondFramesScan = (lambda h: (lambda: h))(g["hbonds"][k].tolist())
            out.append(c)
        return out

    ora = stats_oracle.key_statistics(g["score"], g["hbonds"], 1, 0)
    titles = lambda cs: [c.title for c in cs]
    base = contacts()
    assert titles(cf.TotalTimeFilter("t", u"greater", 25).filterContacts(contacts())) == \
        [c.title for c, v in zip(base, ora["total_time"]) if v > 25]
    assert titles(cf.ScoreFilter("s", u"smaller", 0.3, u"Mean").filterContacts(contacts())) == \
        [c.title for c, v in zip(base, g["mean"]) if v < 0.3]
    assert titles(cf.ScoreFilter("s", u"greater", 0.1, u"Median").filterContacts(contacts())) == \
        [c.title for c, v in zip(base, g["median"]) if v > 0.1]
    assert titles(cf.ScoreFilter("s", u"greater", 10.0, u"HB %").filterContacts(contacts())) == \
        [c.title for c, v in zip(base, g["hbond_percentage"]) if v > 10.0]
    assert titles(cf.OnlyFilter("o", "hbonds", 0).filterContacts(contacts())) == \
        [c.title for c, v in zip(base, g["hbond_percentage"]) if v > 0]
    ns, thr = g["params_b"]
    for key, col in ((u"total time", "total_time_b"), (u"mean lifetime", "mean_life_time_b"),
                     (u"median lifetime", "median_life_time_b")):
        s = cf.Sorting("x", key, True)
        s.setThresholdAndNsPerFrame(thr, ns)
        got = s.sortContacts(contacts())
        vals = [getattr(c, {"total time": "ttime", "mean lifetime": "meanLifeTime", "median lifetime": "medianLifeTime"}[key])
                for c in got]
        assert vals == sorted(vals, reverse=True)
        np.testing.assert_allclose(sorted(vals), np.sort(g[col]), rtol=STATS_RTOL)
