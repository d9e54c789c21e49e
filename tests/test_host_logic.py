"""Host-side logic without a GPU: topology pre-pass, record ordering, accumulation
finalisation, key strings, scheduler slicing, the C-ABI's exported symbols.

Device output is imitated by turning ORACLE results into the record arrays the CUDA
library would return (shuffled, so the host has to restore the reference's order)."""
import os
import pickle
import re

import numpy as np
import pytest

from oracle import frame_oracle
from pycontact_b200 import _lib, prepare, synth
from pycontact_b200.core.Biochemistry import AccumulatedContact, AtomContact, HydrogenBond
from pycontact_b200.core.ContactAnalyzer import Analyzer
from pycontact_b200.engine import BatchResult
from pycontact_b200.results import AccumulationTable, LazyContactResults, TrajectoryContacts
from pycontact_b200.scheduler import chunks, default_batch_frames, prefetched, worker_slices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pycontact_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pc_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "libpycontact_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared
    assert lib.pc_version() >= 100


def test_record_layouts_match_header():
    assert _lib.contact_dtype.itemsize == 16 and _lib.hbond_dtype.itemsize == 20 and _lib.row_dtype.itemsize == 48


def test_worker_slices_assign_every_frame_exactly_once():
    """chunks() keeps the reference's float stepping (num + 1 slices for e.g. 8 frames over 6 workers);
    the scheduler and the rank slicing must still hand every frame to exactly one worker."""
    from pycontact_b200 import distributed
    assert len([c for c in chunks(list(range(8)), 6) if len(c)]) == 7          # the case that used to drop frame 7
    for F in list(range(0, 400)) + [712, 999, 1000, 2999, 3000, 10000]:
        for world in range(1, 17):
            sl = worker_slices(F, world)
            assert len(sl) == world
            assert [f for lo, hi in sl for f in range(lo, hi)] == list(range(F)), (F, world, sl)
            assert [distributed.frame_slice(F, r, world) for r in range(world)] == sl
            ref = [c for c in chunks(list(range(F)), world) if len(c)]
            if len(ref) == world:              # where the reference's rule yields one slice per worker, it is kept
                assert [(c[0], c[-1] + 1) for c in ref] == [x for x in sl if x[1] > x[0]]


def test_prefetched_reader_thread_keeps_order_and_errors():
    """scheduler.prefetched: batches come out in order, a reader's exception reaches the consumer, and a consumer
    that stops early does not leave the reader blocked."""
    import threading
    import time

    def gen(n, fail_at=None):
        for i in range(n):
            if i == fail_at:
                raise ValueError("bad frame %d" % i)
            time.sleep(0.002)
            yield i, np.full(4, i)
    for depth in (0, 1, 3):
        got = list(prefetched(gen(17), depth))
        assert [g[0] for g in got] == list(range(17)) and all((g[1] == g[0]).all() for g in got)
        with pytest.raises(ValueError):
            list(prefetched(gen(9, fail_at=5), depth))
    g = prefetched(gen(1000), 2)
    assert next(g)[0] == 0 and next(g)[0] == 1
    g.close()
    time.sleep(0.3)
    assert not [t for t in threading.enumerate() if t.name == "pycontact-frame-reader" and t.is_alive()]


def test_chunks_rule():
    assert chunks(list(range(50)), 2) == [list(range(25)), list(range(25, 50))]
    assert [len(c) for c in chunks(list(range(50)), 8)] == [6, 6, 6, 7, 6, 6, 6, 7]
    assert [len(c) for c in chunks(list(range(10)), 3)] == [3, 3, 4]
    assert sum(chunks(list(range(7)), 4), []) == list(range(7))
    assert default_batch_frames(3097, 1230) >= 1


def _prep(fx, idx1, idx2):
    t = fx.topology
    hcsr = prepare.hydrogen_csr_from_bond_table(t.names, t.types, t.bonds)
    bb = np.zeros(t.n_atoms, bool)
    bb[fx.backbone] = True
    seg = prepare.string_codes(t.segids)
    s1 = prepare.prepare_selection(idx1, t.names, t.resids, seg, bb, hcsr)
    s2 = s1 if idx2 is None else prepare.prepare_selection(idx2, t.names, t.resids, seg, bb, hcsr)
    return s1, s2, hcsr


def test_hydrogen_csr_matches_oracle(rpn11, golden_ab):
    t = rpn11.topology
    ptr, hyd = prepare.hydrogen_csr_from_bond_table(t.names, t.types, t.bonds)
    optr, ohyd = frame_oracle.bonded_hydrogens(t.names, t.types, t.bonds)
    assert np.array_equal(ptr, optr) and np.array_equal(hyd, ohyd)
    # the ConvBond-shaped input gives the same CSR
    bptr, bid = t.bonds_of_atoms()

    class B:
        def __init__(self, a):
            ids = bid[bptr[a]:bptr[a + 1]]
            self.types = [(t.types[i], t.types[j]) for i, j in t.bonds[ids]]
            self.idx = [t.bonds[b] for b in ids]

        def to_indices(self):
            return self.idx
    p2, h2 = prepare.hydrogen_csr_from_bond_objects([B(a) for a in range(t.n_atoms)], t.names)
    assert np.array_equal(p2, optr) and np.array_equal(h2, ohyd)
    # local CSR: hydrogens outside the selection become -1
    idx1 = golden_ab["sel1"].astype(np.int64)
    s1, _, _ = _prep(rpn11, idx1, None)
    assert (s1.hidx >= 0).all()
    heavy = np.asarray([i for i in idx1 if not t.names[i].startswith("H")], np.int64)
    sh, _, _ = _prep(rpn11, heavy, None)
    assert (sh.hidx == -1).all() and sh.hidx.size > 0
    assert ((s1.flags & _lib.PC_ATOM_HYDROGEN) > 0).sum() == sum(t.names[i].startswith("H") for i in idx1)


def _device_order_keys(fr):
    """Order keys in the device's layout (local index 1 << 37 | stencil rank << 32 | local index 2, pc_order.cuh) whose
    sorted order is the oracle's: the rank of a contact = descents of atom 2's index so far in its atom 1's run."""
    i1, i2 = np.asarray(fr["lidx1"], np.int64), np.asarray(fr["lidx2"], np.int64)
    rank = np.zeros(i1.size, np.uint64)
    for c in range(1, i1.size):
        if i1[c] == i1[c - 1]:
            rank[c] = rank[c - 1] + (1 if i2[c] < i2[c - 1] else 0)
    assert rank.max(initial=0) < 27
    return (i1.astype(np.uint64) << np.uint64(37)) | (rank << np.uint64(32)) | i2.astype(np.uint64)


def _fake_batch(frames, sel1, sel2, rng, frame0=0, with_rows=None, device_keys=False):
    """Oracle frames -> the arrays the device would return, records shuffled inside each frame."""
    loc1 = {int(g): k for k, g in enumerate(sel1.indices)}
    loc2 = {int(g): k for k, g in enumerate(sel2.indices)}
    cs, hs, ks = [], [], []
    ccount, hcount = [], []
    for fr in frames:
        n = len(fr["idx1"])
        rec = np.zeros(n, _lib.contact_dtype)
        rec["i"] = fr["lidx1"]; rec["j"] = fr["lidx2"]
        rec["distance"] = fr["distance"]; rec["weight"] = fr["weight"]
        key = _device_order_keys(fr) if device_keys else np.arange(n, dtype=np.uint64) * np.uint64(7) + np.uint64(3)  # any increasing key
        p = rng.permutation(n)
        cs.append(rec[p]); ks.append(key[p]); ccount.append(n)
        m = len(fr["hb_donor"])
        hb = np.zeros(m, _lib.hbond_dtype)
        c = fr["hb_contact"]
        hb["i"] = fr["lidx1"][c]; hb["j"] = fr["lidx2"][c]
        side = (fr["hb_side"] == 2)
        hl = np.asarray([(loc2 if s else loc1)[int(h)] for s, h in zip(side, fr["hb_hydrogen"])], np.uint32)
        hb["hyd"] = hl | (side.astype(np.uint32) << np.uint32(31))
        hb["distance"] = fr["hb_distance"]; hb["angle"] = fr["hb_angle"]
        hs.append(hb[rng.permutation(m)]); hcount.append(m)
    ccount, hcount = np.asarray(ccount, np.uint32), np.asarray(hcount, np.uint32)
    cstart = (np.cumsum(ccount) - ccount).astype(np.uint32)
    hstart = (np.cumsum(hcount) - hcount).astype(np.uint32)
    return BatchResult(frame0, len(frames), np.concatenate(cs), np.concatenate(ks), np.concatenate(hs), None,
                       cstart, ccount, hstart, hcount, None, None, 0, 1)


def test_trajectory_contacts_restore_reference_order(rpn11, golden_ab):
    idx1, idx2 = golden_ab["sel1"].astype(np.int64), golden_ab["sel2"].astype(np.int64)
    s1, s2, hcsr = _prep(rpn11, idx1, idx2)
    t = rpn11.topology
    frames = [0, 1, 2, 3, 4, 5]
    ora = [frame_oracle.frame_contacts(rpn11.coords[f, idx1], rpn11.coords[f, idx2], idx1, idx2, 5.0, 2.5, 120,
                                       t.names, hcsr) for f in frames]
    rng = np.random.default_rng(0)
    traj = TrajectoryContacts(s1, s2, 2.5, 120)
    traj.add_batch(_fake_batch(ora[:4], s1, s2, rng), frames[:4])
    traj.add_batch(_fake_batch(ora[4:], s1, s2, rng, 4), frames[4:])
    traj.finalize()
    fp = golden_ab["frame_ptr"]
    assert np.array_equal(traj.frame_ptr, fp[:7])
    assert np.array_equal(traj.idx1, golden_ab["idx1"][:fp[6]]) and np.array_equal(traj.idx2, golden_ab["idx2"][:fp[6]])
    hp = golden_ab["hb_ptr"]
    assert np.array_equal(traj.hb_ptr, hp[:fp[6] + 1])
    n = hp[fp[6]]
    for k in ("hb_donor", "hb_acceptor", "hb_hydrogen"):
        assert np.array_equal(getattr(traj, k), golden_ab[k][:n])
    # the packed one-pass sort and the lexsort fallback give the same arrays
    rng = np.random.default_rng(0)
    slow = TrajectoryContacts(s1, s2, 2.5, 120)
    slow._packable = False
    slow.add_batch(_fake_batch(ora[:4], s1, s2, rng), frames[:4])
    slow.add_batch(_fake_batch(ora[4:], s1, s2, rng, 4), frames[4:])
    slow.finalize()
    assert traj._packable is True                    # (these fake order keys are not the device's: sorted as they are)
    # ... and with order keys in the device's layout the one-pass sort itself runs
    rng = np.random.default_rng(1)
    fast = TrajectoryContacts(s1, s2, 2.5, 120)
    fast.add_batch(_fake_batch(ora[:4], s1, s2, rng, device_keys=True), frames[:4])
    fast.add_batch(_fake_batch(ora[4:], s1, s2, rng, 4, device_keys=True), frames[4:])
    fast.finalize()
    assert fast.packed_sorts == 2 and traj.packed_sorts == 0
    for k in ("frame_ptr", "li", "lj", "idx1", "idx2", "distance", "weight", "hb_ptr", "hb_donor", "hb_acceptor",
              "hb_hydrogen", "hb_side", "hb_distance", "hb_angle"):
        assert np.array_equal(getattr(traj, k), getattr(fast, k)), k
    for k in ("frame_ptr", "li", "lj", "idx1", "idx2", "distance", "weight", "hb_ptr", "hb_donor", "hb_acceptor",
              "hb_hydrogen", "hb_side", "hb_distance", "hb_angle"):
        assert np.array_equal(getattr(traj, k), getattr(slow, k)), k
    # object views
    res = LazyContactResults(traj)
    assert len(res) == 6 and len(res[2]) == fp[3] - fp[2]
    c = res[2][5]
    assert isinstance(c, AtomContact) and c.frame == 2 and c.idx1 == golden_ab["idx1"][fp[2] + 5]
    hb = [h for fr in res for c in fr for h in c.hbondinfo]
    assert len(hb) == n and isinstance(hb[0], HydrogenBond) and hb[0].usedCutoffAngle == 120
    assert len(pickle.loads(pickle.dumps(res))) == 6


def _rows_from_oracle(ora, gm1, gm2, sel1, sel2, backbone):
    """(frame, key) cells like pc_accumulate's output, from oracle frames."""
    bb = set(int(b) for b in backbone)
    rows, frames = [], []
    for f, fr in enumerate(ora):
        nhb = np.diff(fr["hb_ptr"])
        cells = {}
        for c in range(len(fr["idx1"])):
            key = int(gm1.gid[fr["lidx1"][c]]) * gm2.n_groups + int(gm2.gid[fr["lidx2"][c]])
            e = cells.setdefault(key, [0.0, 0.0, 0.0, 0, 0, c])
            w = float(fr["weight"][c])
            e[0] += w
            e[1] += w if int(fr["idx1"][c]) in bb else 0.0
            e[2] += w if int(fr["idx2"][c]) in bb else 0.0
            e[3] += 1
            e[4] += int(nhb[c])
        for key, e in cells.items():
            rows.append((key, e[0], e[1], e[2], e[3], e[4], e[5]))
            frames.append(f)
    return np.asarray(rows, _lib.row_dtype), np.asarray(frames, np.int64)


def test_accumulation_table_matches_golden(rpn11, golden_ab):
    idx1, idx2 = golden_ab["sel1"].astype(np.int64), golden_ab["sel2"].astype(np.int64)
    s1, s2, hcsr = _prep(rpn11, idx1, idx2)
    t = rpn11.topology
    ora = [frame_oracle.frame_contacts(rpn11.coords[f, idx1], rpn11.coords[f, idx2], idx1, idx2, 5.0, 2.5, 120,
                                       t.names, hcsr) for f in range(50)]
    gm1 = prepare.group_map(idx1, golden_ab["map1"], t.names, t.resids, t.resnames, t.segids)
    gm2 = prepare.group_map(idx2, golden_ab["map2"], t.names, t.resids, t.resnames, t.segids)
    assert gm1.n_groups == len({(t.resids[i], t.resnames[i]) for i in idx1})
    rows, frames = _rows_from_oracle(ora, gm1, gm2, s1, s2, rpn11.backbone)
    p = np.random.default_rng(1).permutation(len(rows))
    table = AccumulationTable(50, gm1, gm2)
    table.add_rows(rows[p], frames[p])
    table.finalize()
    assert table.n_keys == 148
    objs = table.to_objects(t.names, t.resids, t.resnames, t.segids)
    assert [o.key1 for o in objs] == golden_ab["acc_key1"].tolist()
    assert [o.key2 for o in objs] == golden_ab["acc_key2"].tolist()
    assert [o.title for o in objs] == golden_ab["acc_title"].tolist()
    np.testing.assert_allclose(table.score_matrix(), golden_ab["acc_score"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(np.asarray([o.scoreArray for o in objs], float), golden_ab["acc_score"], rtol=1e-12)
    assert np.array_equal(table.hbond_matrix(), golden_ab["acc_hbonds"])
    for k in ("bb1", "sc1", "bb2", "sc2"):
        np.testing.assert_allclose(getattr(table, k), golden_ab["acc_" + k], rtol=1e-9, atol=1e-9)
    assert sum(o.hbond_percentage() for o in objs) == 676.0              # tests/test_basic.py:43
    assert table.hbond_percentage().sum() == 676.0
    assert objs[0].scoreArray.count(0) == int((golden_ab["acc_score"][0] == 0).sum())
    o2 = pickle.loads(pickle.dumps(objs[3]))
    assert o2.scoreArray == objs[3].scoreArray and o2.key1 == objs[3].key1 and o2.hbondFrames == objs[3].hbondFramesScan()


def _finalize_reference(rows, frames):
    """AccumulationTable.finalize, the plain way: np.unique + lexsorts (what the vectorised version replaced)."""
    keys, kid = np.unique(rows["key"], return_inverse=True)
    K = keys.size
    o = np.lexsort((rows["first"], frames))
    _, first_pos = np.unique(kid[o], return_index=True)
    rank_of_key = np.argsort(np.argsort(first_pos, kind="stable"), kind="stable")
    kpos = rank_of_key[kid]
    o2 = np.lexsort((frames, kpos))
    out_keys = np.empty(K, np.uint64)
    out_keys[rank_of_key] = keys
    key_ptr = np.zeros(K + 1, np.int64)
    np.cumsum(np.bincount(kpos, minlength=K), out=key_ptr[1:])
    return out_keys, key_ptr, rows[o2], frames[o2]


@pytest.mark.parametrize("n_rows,n_keys,n_frames,seed", [(0, 0, 5, 0), (1, 1, 1, 1), (50, 1, 50, 2), (400, 37, 60, 3),
                                                         (5000, 300, 700, 4), (3000, 3000, 1, 5)])
def test_accumulation_table_finalize_equals_plain_version(n_rows, n_keys, n_frames, seed):
    """Keys in first-appearance order (smallest frame, then smallest order key), cells by (key, frame): the two-argsort
    finalize against np.unique + lexsort on random cells -- one cell per (frame, key), frames in any order over batches."""
    rng = np.random.default_rng(seed)

    class GM:
        n_groups = 1 << 20
    cells = rng.choice(max(n_keys * n_frames, 1), n_rows, replace=False) if n_rows else np.zeros(0, np.int64)
    key_values = rng.choice(1 << 40, max(n_keys, 1), replace=False).astype(np.uint64)
    rows = np.zeros(n_rows, _lib.row_dtype)
    rows["key"] = key_values[cells % max(n_keys, 1)] if n_rows else 0
    frames = (cells // max(n_keys, 1)).astype(np.int64) + 3            # frame positions need not start at 0
    rows["score"] = rng.random(n_rows); rows["bb1"] = rows["score"] * 0.25; rows["bb2"] = rows["score"] * 0.5
    rows["ncontacts"] = rng.integers(1, 9, n_rows); rows["nhbonds"] = rng.integers(0, 3, n_rows)
    rows["first"] = rng.choice(1 << 50, max(n_rows, 1), replace=False).astype(np.uint64)[:n_rows]
    perm = rng.permutation(n_rows)
    t = AccumulationTable(n_frames + 3, GM(), GM())
    for part in np.array_split(perm, 3):
        t.add_rows(rows[part], frames[part])
    t.finalize()
    keys, key_ptr, srows, sframes = _finalize_reference(rows[perm], frames[perm])
    assert np.array_equal(t.keys, keys) and np.array_equal(t.key_ptr, key_ptr)
    assert np.array_equal(t.row_frames, sframes) and np.array_equal(t.rows, srows)
    assert np.array_equal(t.row_key, np.repeat(np.arange(keys.size), np.diff(key_ptr)))
    np.testing.assert_allclose(t.total_score, [srows["score"][a:b].sum() for a, b in zip(key_ptr[:-1], key_ptr[1:])], rtol=1e-12)


def test_key_helpers_round_trip():
    an = Analyzer("a.psf", "a.dcd", 5.0, 2.5, 120, "segid A", "segid B")
    an.setTrajectoryData(["ASP", "LYS"], [21, 11], ["OD1", "NZ"], ["RN11", "UBQ"], [], "segid A", "segid B")
    c = AtomContact(0, 3.0, 0.9, 0, 1, [])
    k1, k2 = an.makeKeyArraysFromMaps([0, 0, 1, 1, 0], [0, 0, 1, 1, 0], c)
    assert k1 == ["none", "none", 21, "ASP", "none"] and k2 == ["none", "none", 11, "LYS", "none"]
    key = Analyzer.makeKeyFromKeyArrays(k1, k2)
    assert key == "r.21rn.ASP-r.11rn.LYS"
    assert an.makeKeyArraysFromKey(key) == [["none", "none", "21", "ASP", "none"], ["none", "none", "11", "LYS", "none"]]
    k1, k2 = an.makeKeyArraysFromMaps([1, 1, 1, 1, 1], [1, 0, 0, 0, 1], c)
    key = Analyzer.makeKeyFromKeyArrays(k1, k2)
    assert key == "i.0nm.OD1r.21rn.ASPs.RN11-i.1s.UBQ"
    assert an.makeKeyArraysFromKey(key) == [["0", "OD1", "21", "ASP", "RN11"], ["1", "none", "none", "none", "UBQ"]]
    acc = AccumulatedContact(["none", "none", "21", "ASP", "none"], ["none", "none", "11", "LYS", "none"])
    assert acc.title == "ASP21 - LYS11"
    assert Analyzer.make_single_title(["5", "CA", "21", "ASP", "A"]) == "ASP21, i. 5, nm. CA, s. A"


def test_accumulated_contact_statistics():
    acc = AccumulatedContact(["none"] * 5, ["none"] * 5)
    for s in [0, 1.0, 2.0, 0.2, 0, 3.0, 3.0]:
        acc.addScore(s)
    acc.addContributingAtoms([])
    assert acc.total_time(2, 0.5) == 8
    assert acc.life_time(1, 0.5) == [2, 2]
    assert acc.life_time(1, 2.5) == [2]
    acc.scoreArray = [1.0, 0, 0]
    assert acc.life_time(1, 0.5) == [1, 0]
    assert acc.mean_score() == pytest.approx(1 / 3) and acc.median_score() == 0
    acc.bb1, acc.sc1, acc.bb2, acc.sc2 = 2.0, 1.0, 0.0, 3.0
    assert acc.determineBackboneSidechainType() == 1
    salt = AccumulatedContact(["none", "none", "1", "ASP", "none"], ["none", "none", "2", "LYS", "none"])
    salt.scoreArray = [1.0]
    salt.contributingAtoms = [[]]
    salt.sc1 = salt.sc2 = 1.0
    assert salt.contactTypeAsShortcut() == "saltbr"


def test_selection_language_and_empty_selection(rpn11):
    from pycontact_b200.io.selection import select_atoms
    t = rpn11.topology
    assert len(select_atoms(t, "segid RN11")) == 3097 and len(select_atoms(t, "segid UBQ")) == 1230
    assert np.array_equal(select_atoms(t, "backbone"), rpn11.backbone)
    assert len(select_atoms(t, "segid A")) == 0


def test_synthetic_systems_have_named_sizes():
    s = synth.make_system("tiny")
    assert s.topology.n_atoms == 1128 and len(s.sel1) == len(s.sel2) == 24 * 16
    c = synth.frames_host(s, [0, 3])
    assert c.dtype == np.float32 and c.shape == (2, 1128, 3)
