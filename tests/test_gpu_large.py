"""Large-frame path and accumulation WITHOUT order keys vs the oracle -- needs a B200 (``-m gpu``).

These are the kernels the 1 M-atom (c3) and 3 M-atom (c4) bench lines time: the fused per-frame tail
kernel (pc_scan_tail.cuh), the compact shared-memory accumulation (pc_compact_frame), the global
tile-table accumulation without order keys, and the dense pair kernel.  Without order keys a frame's
contacts come back in device order, so frames are compared as SETS: pairs and H-bond triples bit-exact,
distances / weights / angles / per-cell scores within 1e-5 relative (BASELINE.json north_star).
Everything goes through the C ABI (ctypes -> libpycontact_b200.so).
"""
import numpy as np
import pytest

from oracle import frame_oracle, native
from pycontact_b200 import prepare, synth
from pycontact_b200.engine import ContactEngine
from pycontact_b200.scheduler import ArraySource, run_scan
from tests.test_gpu_parity import RTOL, _selections

pytestmark = pytest.mark.gpu


def _compare_unordered(traj, oracle_frames):
    """TrajectoryContacts without order keys lists a frame's contacts by (idx1, idx2)."""
    assert traj.n_frames == len(oracle_frames)
    for f, fr in enumerate(oracle_frames):
        lo, hi = int(traj.frame_ptr[f]), int(traj.frame_ptr[f + 1])
        p = np.lexsort((fr["idx2"], fr["idx1"]))
        assert hi - lo == p.size, "frame %d: %d contacts, oracle %d" % (f, hi - lo, p.size)
        assert np.array_equal(traj.idx1[lo:hi], fr["idx1"][p]), "frame %d: pair set differs" % f
        assert np.array_equal(traj.idx2[lo:hi], fr["idx2"][p]), "frame %d: pair set differs" % f
        np.testing.assert_allclose(traj.distance[lo:hi], fr["distance"][p], rtol=RTOL)
        np.testing.assert_allclose(traj.weight[lo:hi], fr["weight"][p], rtol=RTOL, atol=1e-30)
        # H-bonds regrouped by contact in the permuted order (inside a contact: atom 1's hydrogens, then atom 2's)
        cnt = np.diff(fr["hb_ptr"])[p]
        src = np.repeat(fr["hb_ptr"][:-1][p], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        hlo, hhi = int(traj.hb_ptr[lo]), int(traj.hb_ptr[hi])
        assert np.array_equal(np.diff(traj.hb_ptr[lo:hi + 1]), cnt), "frame %d: H-bond flags differ" % f
        for k in ("hb_donor", "hb_acceptor", "hb_hydrogen"):
            assert np.array_equal(getattr(traj, k)[hlo:hhi], fr[k][src]), "frame %d: %s differs" % (f, k)
        np.testing.assert_allclose(traj.hb_distance[hlo:hhi], fr["hb_distance"][src], rtol=RTOL)
        np.testing.assert_allclose(traj.hb_angle[hlo:hhi], fr["hb_angle"][src], rtol=RTOL)


def _compare_cells(table, oracle_cells, code1, code2):
    """(frame, key) cells of the device vs accumulate_arrays: key sets exact, sums within RTOL, counts exact."""
    k1 = code1[table.gm1.rep[table.g1[table.row_key]]]
    k2 = code2[table.gm2.rep[table.g2[table.row_key]]]
    for f, cells in enumerate(oracle_cells):
        m = np.nonzero(table.row_frames == f)[0]
        o = m[np.lexsort((k2[m], k1[m]))]
        assert o.size == cells["score"].size, "frame %d: %d cells, oracle %d" % (f, o.size, cells["score"].size)
        assert np.array_equal(k1[o], cells["k1"]) and np.array_equal(k2[o], cells["k2"]), "frame %d: key set differs" % f
        r = table.rows[o]
        np.testing.assert_allclose(r["score"], cells["score"], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(r["bb1"], cells["bb1"], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(r["bb2"], cells["bb2"], rtol=RTOL, atol=1e-9)
        assert np.array_equal(r["ncontacts"], cells["ncont"]) and np.array_equal(r["nhbonds"], cells["nhb"])


def _run(s, coords, self_mode, maps, *, path=2, tail=True, table=0, dense=0, order_keys=False, batch=3,
         cutoff=5.0, search=None, group=False):
    """Scan + accumulate `coords` (all atoms of the system) on the device and through the oracle."""
    t = s.topology
    nf = coords.shape[0]
    idx1 = s.sel1
    idx2 = s.sel1 if self_mode else s.sel2
    x2 = [coords[k, idx2] for k in range(nf)]
    ora = frame_oracle.scan_frames([coords[k, idx1] for k in range(nf)], x2, idx1, idx2, cutoff, 2.5, 120,
                                   t.names, t.types, t.bonds, t.resids, t.segids, self_mode, search)
    s1, s2 = _selections(t, idx1, None if self_mode else idx2, s.backbone)
    gm1 = prepare.group_map(idx1, maps[0], t.names, t.resids, t.resnames, t.segids)
    gm2 = prepare.group_map(idx2, maps[1], t.names, t.resids, t.resnames, t.segids)
    with ContactEngine(s1, s2, cutoff, 2.5, 120) as eng:
        eng.set_path(path)
        eng.set_tail(tail)
        eng.set_dense_mode(dense)
        eng.set_group_mode(group)
        if table:
            eng.set_limits(table=table)
        eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
        src = ArraySource(np.ascontiguousarray(coords[:, idx1]), None if self_mode else np.ascontiguousarray(coords[:, idx2]))
        traj, tab, stats = run_scan([eng], src, batch_frames=batch, order_keys=order_keys, accumulate=True,
                                    group_maps=(gm1, gm2))
        final = dict(tail=eng.tail, lim=list(eng.lim), path=stats["path"], group=any(eng.last_group(k) for k in range(len(eng.slots))))
    c1 = frame_oracle.atom_key_codes(maps[0], t.n_atoms, t.names, t.resids, t.resnames, t.segids)
    c2 = frame_oracle.atom_key_codes(maps[1], t.n_atoms, t.names, t.resids, t.resnames, t.segids)
    cells = frame_oracle.accumulate_arrays(ora, c1, c2, s.backbone, t.n_atoms)
    return traj, tab, ora, cells, c1, c2, final


# tail: fused per-frame kernel or multi-kernel sequence; table: pc_set_limits' table_slots (0 = compact
# shared-memory table, 16384 = global tile table)
@pytest.mark.parametrize("name,self_mode,nframes,tail,table", [
    ("c3_small", False, 5, True, 0),         # fused tail kernel, A from the frame (tail mode 3, default) incl. pc_compact_frame
    ("c3_small", False, 5, 1, 0),            # fused tail kernel on the compacted atoms of both selections (tail mode 1)
    ("c3_small", False, 5, 1, 0),            # tail kernel on the compacted atoms of both selections (tail mode 1)
    ("small", False, 7, 1, 0), ("small_i", False, 4, 1, 0), ("tiny", False, 6, 1, 16384),
    ("c3_small", False, 5, False, 0),        # multi-kernel sequence + pc_accumulate_compact_kernel
    ("c3_small", False, 5, True, 16384),     # fused tail kernel, rows from the global tile table
    ("c3_small", False, 5, False, 16384),
    ("small", False, 7, True, 0), ("small", False, 7, False, 0), ("small", False, 7, False, 16384),
    ("small_i", False, 4, True, 0), ("small_i", False, 4, 1, 16384), ("small_i", False, 4, False, 0),   # interdigitated chains
    ("tiny", False, 6, True, 0), ("tiny", True, 6, False, 0), ("tiny", True, 6, False, 16384),
    ("c4_small", True, 2, False, 0),         # more keys per frame than the compact table: overflow -> global table
    ("c4_small", True, 2, False, 16384),
])
def test_large_path_accumulation_without_order_keys(name, self_mode, nframes, tail, table):
    s = synth.make_system(name)
    coords = synth.frames_host(s, list(range(nframes)))
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, self_mode, s.maps, tail=tail, table=table)
    assert final["path"] == 2
    if not self_mode:
        assert final["tail"] == (3 if tail is True else int(tail))     # a fused kernel did not silently hand the batch back
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


# The group kernel (pc_scan_group.cuh): frames the tail kernel does not take -- self contacts, or tail mode 0 -- searched,
# recorded and accumulated unit by unit of whole groups.  c4_small is the scaled capsid of the c4 bench line.
@pytest.mark.parametrize("name,self_mode,nframes,batch", [
    ("c4_small", True, 3, 2), ("tiny", True, 6, 3), ("small", True, 4, 3),
    ("c3_small", False, 5, 3), ("small", False, 7, 3), ("small_i", False, 4, 4), ("tiny", False, 6, 1),
])
def test_group_kernel_matches_oracle(name, self_mode, nframes, batch):
    s = synth.make_system(name)
    coords = synth.frames_host(s, list(range(nframes)))
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, self_mode, s.maps, tail=False, group=True, batch=batch)
    assert final["path"] == 2 and final["group"], "the group kernel did not run"
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


def test_group_kernel_is_skipped_when_groups_are_not_runs():
    """Atom-name maps: a group's atoms are scattered over the selection, so no unit sees all contacts of a key --
    the scan falls back to the multi-kernel sequence + table; results unchanged."""
    s = synth.make_system("small")
    coords = synth.frames_host(s, [0, 1, 2])
    maps = ([0, 1, 0, 0, 0], [0, 1, 0, 0, 0])
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, False, maps, tail=False, group=True)
    assert not final["group"]
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


def test_group_kernel_table_overflow_falls_back():
    """Per-atom keys on selection 2 at a 6.5 A cutoff: more than 1024 keys per unit -> PC_ST_TABLE_OVERFLOW -> the engine
    re-runs the batch on the global table."""
    s = synth.make_system("c4_small")
    coords = synth.frames_host(s, [0])
    maps = ([0, 0, 1, 1, 1], [1, 0, 0, 0, 0])
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, True, maps, tail=False, group=True, batch=1, cutoff=6.5)
    assert final["lim"][3] == 16384 and not final["group"]
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


def test_fused_tail_hands_over_frames_that_do_not_fit():
    """More in-grid B atoms than the tail kernel's shared-memory pool: PC_ST_TAIL_OVERFLOW -> the engine re-runs
    the batch on the multi-kernel sequence; results unchanged."""
    a, b, ha, hb = synth.cloud(6000, 60000, seed=21)
    s = _cloud_system(a, b, ha, hb)
    coords = s.coords0[None]
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, False, s.maps, tail=True, batch=1)
    assert final["tail"] == 0
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


@pytest.mark.parametrize("path", [1, 2])
def test_coarse_maps_leave_the_float32_tables(path):
    """Segment-level maps: one (frame, key) cell holds every contact of the frame.  The float32 cell sums of the
    fused / compact tables are not trusted beyond PC_FLOAT_SUM_MAX contacts (PC_ST_PRECISION): the engine redoes
    the batch on the exact global table; sums within 1e-5 of the float64 oracle."""
    s = synth.make_system("c2")                 # two chains of one segment each: ~2000 contacts in ONE cell per frame
    coords = synth.frames_host(s, [0, 1, 2])
    maps = ([0, 0, 0, 0, 1], [0, 0, 0, 0, 1])
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, False, maps, path=path)
    assert max(int(c["ncont"].max()) for c in cells) > 1024
    assert final["lim"][3] == 16384
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


def _cloud_system(a, b, heavy_a, heavy_b):
    """Two atom clouds as a system: heavy atoms are carbons / oxygens, the rest hydrogens without bonds."""
    from pycontact_b200.io.psf import Topology
    n1, n2 = len(a), len(b)
    n = n1 + n2
    heavy = np.concatenate([heavy_a, heavy_b])
    names = np.where(heavy, np.where(np.arange(n) % 3 == 0, "O", "CA"), "H1").tolist()
    topo = Topology(["A"] * n1 + ["B"] * n2, (np.arange(n) // 8 + 1).astype(np.int64), ["ALA"] * n, names,
                    ["C"] * n, np.zeros(n), np.ones(n), np.zeros((0, 2), np.int64))
    xyz = np.concatenate([a, b]).astype(np.float32)
    z = np.zeros((1, 3), np.float32)
    return synth.SynthSystem("cloud", 0, topo, xyz, np.zeros(n, np.int32), z, np.ones(1, np.float32), np.zeros(1, np.float32),
                             np.arange(n1, dtype=np.int64), np.arange(n1, n, dtype=np.int64), "segid A", "segid B",
                             np.zeros(0, np.int64), maps=([0, 0, 1, 0, 0], [0, 0, 1, 0, 0]))


@pytest.mark.parametrize("dense", [1, 2])
@pytest.mark.parametrize("n1,n2,cutoff", [(20000, 20000, 5.0), (3000, 40000, 8.0)])
def test_dense_pair_kernels_on_uniform_clouds(dense, n1, n2, cutoff):
    """C5's uniform 0.1 A^-3 clouds through both pair kernels of the multi-kernel sequence (dense = 2 forces
    pc_large_pair_dense_kernel, 1 the thread-per-atom kernel) vs the compiled reference search."""
    a, b, ha, hb = synth.cloud(n1, n2, seed=31)
    s = _cloud_system(a, b, ha, hb)
    coords = np.stack([s.coords0, s.coords0 + np.float32(0.375)])
    search = native.ref_find_contacts if native.have_reference() else None
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, False, s.maps, tail=False, dense=dense, table=16384,
                                                cutoff=cutoff, search=search, batch=2)
    assert sum(len(f["idx1"]) for f in ora) > 100000
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


def test_dense_pair_kernel_self_mode():
    s = synth.make_system("c4_small")
    coords = synth.frames_host(s, [0, 3])
    traj, tab, ora, cells, c1, c2, final = _run(s, coords, True, s.maps, tail=False, dense=2, table=16384, batch=2)
    _compare_unordered(traj, ora)
    _compare_cells(tab, cells, c1, c2)


# ---------------------------------------------------------------------------------------------------
# FULL-SIZE bench systems: frames generated on the device (what bench.py analyses), sampled frames
# copied back and run through the oracle with the compiled reference search.
# ---------------------------------------------------------------------------------------------------
def _sampled(F, seed):
    rng = np.random.default_rng(seed)
    return sorted(set([0, F // 2, F - 1] + rng.integers(1, F - 1, 5).tolist()))


@pytest.mark.parametrize("name,nextra", [("c2", 5), ("c3", 5)])
def test_full_size_bench_systems_on_sampled_frames(name, nextra):
    from pycontact_b200 import device_synth
    s = synth.make_system(name)
    t = s.topology
    frames = _sampled(s.n_frames, 7)
    # the two generators bench.py uses (noise keyed by selection): exactly the bytes its scan reads
    h1 = device_synth.DeviceTrajectory(s, 0, s.sel1).frames_to_host(frames)
    h2 = device_synth.DeviceTrajectory(s, 0, s.sel2).frames_to_host(frames)
    search = native.ref_find_contacts if native.have_reference() else None
    ora = frame_oracle.scan_frames([h1[k] for k in range(len(frames))], [h2[k] for k in range(len(frames))],
                                   s.sel1, s.sel2, 5.0, 2.5, 120, t.names, t.types, t.bonds, search=search)
    s1, s2 = _selections(t, s.sel1, s.sel2, s.backbone)
    gm1 = prepare.group_map(s.sel1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
    gm2 = prepare.group_map(s.sel2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
    c1 = frame_oracle.atom_key_codes(s.maps[0], t.n_atoms, t.names, t.resids, t.resnames, t.segids)
    cells = frame_oracle.accumulate_arrays(ora, c1, c1, s.backbone, t.n_atoms)
    src = ArraySource(h1, h2)
    # the throughput configuration (no order keys: fused accumulation) and the object API's (order keys)
    for order_keys in (False, True):
        with ContactEngine(s1, s2, 5.0, 2.5, 120) as eng:
            eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
            traj, tab, stats = run_scan([eng], src, batch_frames=4, order_keys=order_keys, accumulate=True,
                                        group_maps=(gm1, gm2))
            assert stats["path"] == (1 if name == "c2" else 2)
            if name == "c3":
                assert eng.tail, "the 1 M-atom frames must fit the fused tail kernel"
        if order_keys:
            from tests.test_gpu_parity import _compare_traj
            _compare_traj(traj, ora)                        # the reference's in-frame ORDER too
        else:
            _compare_unordered(traj, ora)
        _compare_cells(tab, cells, c1, c1)
    assert sum(len(f["idx1"]) for f in ora) > 1000 * len(frames)
