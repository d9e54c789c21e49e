"""The oracle (oracle/) against the reference's golden vectors -- CPU only.

Golden data: tests/golden/*.npz, produced by tests/golden/make_golden.py from the
unmodified reference and its pickled session."""
import numpy as np
import pytest

from oracle import frame_oracle, native
from pycontact_b200 import synth


def _scan(fx, sel1, sel2, frames, self_mode, search=None):
    t = fx.topology
    return frame_oracle.scan_frames(
        [fx.coords[f, sel1] for f in frames], [fx.coords[f, sel2] for f in frames], sel1, sel2,
        5.0, 2.5, 120, t.names, t.types, t.bonds, t.resids, t.segids, self_mode, search)


def _check_frames(res, g, prefix="", frames=None):
    fp = g[prefix + "frame_ptr"]
    for k, fr in enumerate(res):
        lo, hi = fp[k], fp[k + 1]
        assert np.array_equal(fr["idx1"], g[prefix + "idx1"][lo:hi])          # same pairs, same ORDER
        assert np.array_equal(fr["idx2"], g[prefix + "idx2"][lo:hi])
        np.testing.assert_allclose(fr["distance"], g[prefix + "distance"][lo:hi], rtol=1e-12)
        np.testing.assert_allclose(fr["weight"], g[prefix + "weight"][lo:hi], rtol=1e-12)
        hp = g[prefix + "hb_ptr"]
        hlo, hhi = hp[lo], hp[hi]
        assert np.array_equal(fr["hb_ptr"], hp[lo:hi + 1] - hlo)
        assert np.array_equal(fr["hb_donor"], g[prefix + "hb_donor"][hlo:hhi])
        assert np.array_equal(fr["hb_acceptor"], g[prefix + "hb_acceptor"][hlo:hhi])
        assert np.array_equal(fr["hb_hydrogen"], g[prefix + "hb_hydrogen"][hlo:hhi])
        np.testing.assert_allclose(fr["hb_distance"], g[prefix + "hb_distance"][hlo:hhi], rtol=1e-12)
        np.testing.assert_allclose(fr["hb_angle"], g[prefix + "hb_angle"][hlo:hhi], rtol=1e-10)


def test_search_oracle_matches_compiled_reference_on_fixture(rpn11, golden_ab):
    if not native.have_reference():
        pytest.skip("oracle/_ref not built (no reference tree in this environment)")
    s1, s2 = golden_ab["sel1"], golden_ab["sel2"]
    for f in (0, 17, 49):
        a = native.oracle_find_contacts(rpn11.coords[f, s1], rpn11.coords[f, s2], 5.0)
        b = native.ref_find_contacts(rpn11.coords[f, s1], rpn11.coords[f, s2], 5.0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert int(a[0][-1]) == int(golden_ab["n_raw"][f])


@pytest.mark.parametrize("n1,n2,cutoff,seed", [(500, 700, 5.0, 1), (2000, 300, 4.6, 2), (1500, 1500, 7.3, 3),
                                               (50, 4000, 3.0, 4), (1, 1, 5.0, 5), (300, 300, 12.0, 6)])
def test_search_oracle_matches_compiled_reference_on_clouds(n1, n2, cutoff, seed):
    if not native.have_reference():
        pytest.skip("oracle/_ref not built")
    a, b, _, _ = synth.cloud(n1, n2, seed=seed)
    ra = native.oracle_find_contacts(a, b, cutoff)
    rb = native.ref_find_contacts(a, b, cutoff)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])      # incl. in-row order
    tot, counts = native.brute_count(a, b, cutoff)                            # set == brute force
    assert tot == int(ra[0][-1]) and np.array_equal(counts, np.diff(ra[0]))


def test_search_oracle_self_and_empty():
    a, _, _, _ = synth.cloud(400, 1, seed=9)
    rp, cols = native.oracle_find_contacts(a, a, 5.0)
    rows = native.as_list_of_lists(rp, cols, 400)
    assert len(rows) == 800 and all(len(r) == 0 for r in rows[400:])          # Q3
    assert all(i in rows[i] for i in range(400))                              # (i,i) pairs kept, Q2
    rp, cols = native.oracle_find_contacts(np.zeros((0, 3), np.float32), a, 5.0)
    assert rp.tolist() == [0] and cols.size == 0


def test_canonical_order_is_idx1_rank_idx2():
    """App. A.3: a row's neighbours are ordered by (stencil rank, idx2)."""
    a, b, _, _ = synth.cloud(800, 900, seed=12)
    rp, cols = native.oracle_find_contacts(a, b, 5.0)
    gmin, inv, nb = native.grid_geometry(a, b, 5.0)
    ca, cb = native.box_coords(a, gmin, inv, nb), native.box_coords(b, gmin, inv, nb)
    rank = native.stencil_rank_table()
    for i in range(0, 800, 37):
        js = cols[rp[i]:rp[i + 1]]
        d = cb[js] - ca[i]
        r = rank[d[:, 0] + 1, d[:, 1] + 1, d[:, 2] + 1]
        key = list(zip(r.tolist(), js.tolist()))
        assert key == sorted(key) and (r >= 0).all()


def test_frame_oracle_reproduces_golden_session(rpn11, golden_ab):
    s1, s2 = golden_ab["sel1"].astype(np.int64), golden_ab["sel2"].astype(np.int64)
    res = _scan(rpn11, s1, s2, range(50), False)
    assert sum(len(r["idx1"]) for r in res) == 23208
    assert sum(len(r["hb_donor"]) for r in res) == 457
    assert [r["n_raw"] for r in res] == golden_ab["n_raw"].tolist()
    _check_frames(res, golden_ab)
    t = rpn11.topology
    acc = frame_oracle.accumulate(res, golden_ab["map1"], golden_ab["map2"], t.names, t.resids, t.resnames,
                                  t.segids, rpn11.backbone)
    assert len(acc["keys"]) == 148                                             # tests/test_basic.py:39
    assert np.array_equal(np.asarray(acc["key1"]), golden_ab["acc_key1"])      # first-appearance order
    assert np.array_equal(np.asarray(acc["key2"]), golden_ab["acc_key2"])
    np.testing.assert_allclose(acc["score"], golden_ab["acc_score"], rtol=1e-12, atol=1e-300)
    for k in ("bb1", "sc1", "bb2", "sc2"):
        np.testing.assert_allclose(acc[k], golden_ab["acc_" + k], rtol=1e-12, atol=1e-12)
    assert np.array_equal(acc["hbonds"], golden_ab["acc_hbonds"])
    assert np.array_equal(acc["ncontrib"], golden_ab["acc_ncontrib"])
    assert frame_oracle.hbond_percentage(acc).sum() == 676.0                   # tests/test_basic.py:43


def test_array_accumulation_oracle_equals_the_loop_oracle(rpn11, golden_ab):
    """accumulate_arrays (what the full-size GPU tests use) gives the cells of the reference's golden session."""
    s1, s2 = golden_ab["sel1"].astype(np.int64), golden_ab["sel2"].astype(np.int64)
    res = _scan(rpn11, s1, s2, range(50), False)
    t = rpn11.topology
    for m1, m2 in ((golden_ab["map1"], golden_ab["map2"]), ([0, 1, 1, 1, 1], [0, 0, 0, 0, 1]), ([1, 0, 0, 0, 0], [0, 0, 1, 0, 0])):
        acc = frame_oracle.accumulate(res, m1, m2, t.names, t.resids, t.resnames, t.segids, rpn11.backbone)
        c1 = frame_oracle.atom_key_codes(m1, t.n_atoms, t.names, t.resids, t.resnames, t.segids)
        c2 = frame_oracle.atom_key_codes(m2, t.n_atoms, t.names, t.resids, t.resnames, t.segids)
        arr = frame_oracle.accumulate_arrays(res, c1, c2, rpn11.backbone, t.n_atoms)
        # code pair -> key strings through any contact that carries it
        name_of = {}
        for fr in res:
            for a, b in zip(fr["idx1"].tolist(), fr["idx2"].tolist()):
                name_of.setdefault((int(c1[a]), int(c2[b])), frame_oracle._key_string(
                    frame_oracle._key_arrays([int(bool(v)) for v in m1], a, t.names, t.resids, t.resnames, t.segids),
                    frame_oracle._key_arrays([int(bool(v)) for v in m2], b, t.names, t.resids, t.resnames, t.segids)))
        row = {k: i for i, k in enumerate(acc["keys"])}
        assert len(set(name_of.values())) == len(name_of) == len(acc["keys"])
        bb1 = np.zeros(len(row))
        for f, fr in enumerate(arr):
            ks = [row[name_of[(int(a), int(b))]] for a, b in zip(fr["k1"], fr["k2"])]
            assert sorted(ks) == np.nonzero(acc["ncontrib"][:, f])[0].tolist()
            np.testing.assert_allclose(fr["score"], acc["score"][ks, f], rtol=1e-12)
            assert np.array_equal(fr["nhb"], acc["hbonds"][ks, f]) and np.array_equal(fr["ncont"], acc["ncontrib"][ks, f])
            np.add.at(bb1, ks, fr["bb1"])
        np.testing.assert_allclose(bb1, acc["bb1"], rtol=1e-12, atol=1e-12)


def test_frame_oracle_self_mode(rpn11, golden_self):
    s1 = golden_self["sel1"].astype(np.int64)
    frames = golden_self["full_frames"].tolist()
    res = _scan(rpn11, s1, s1, frames, True)
    _check_frames(res, golden_self, "full_")
    for k, f in enumerate(frames):
        assert len(res[k]["idx1"]) == golden_self["frame_counts"][f]
        assert len(res[k]["hb_donor"]) == golden_self["frame_hbonds"][f]
        assert (res[k]["idx1"] != res[k]["idx2"]).all()
    # whole trajectory: counts, H-bonds, weight sums, accumulation
    res = _scan(rpn11, s1, s1, range(50), True)
    assert [len(r["idx1"]) for r in res] == golden_self["frame_counts"].tolist()
    assert [len(r["hb_donor"]) for r in res] == golden_self["frame_hbonds"].tolist()
    np.testing.assert_allclose([r["weight"].sum() for r in res], golden_self["frame_weight_sum"], rtol=1e-12)
    t = rpn11.topology
    acc = frame_oracle.accumulate(res, [0, 0, 1, 1, 0], [0, 0, 1, 1, 0], t.names, t.resids, t.resnames,
                                  t.segids, rpn11.backbone)
    assert len(acc["keys"]) == 788
    assert np.array_equal(np.asarray(acc["key1"]), golden_self["acc_key1"])
    np.testing.assert_allclose(acc["score"], golden_self["acc_score"], rtol=1e-12, atol=1e-300)
    assert frame_oracle.hbond_percentage(acc).sum() == 5212.0


@pytest.mark.parametrize("tag,self_mode", [("ab", False), ("self", True)])
def test_frame_oracle_synthetic(golden_synth, tag, self_mode):
    s = synth.make_system("tiny", seed=7)
    coords = golden_synth["coords"]
    assert np.array_equal(coords, synth.frames_host(s, range(6)))              # generator is deterministic
    t = s.topology
    s2 = s.sel1 if self_mode else s.sel2
    res = frame_oracle.scan_frames([coords[f, s.sel1] for f in range(6)], [coords[f, s2] for f in range(6)],
                                   s.sel1, s2, 5.0, 2.5, 120, t.names, t.types, t.bonds, t.resids, t.segids,
                                   self_mode)
    _check_frames(res, golden_synth, tag + "_")
    acc = frame_oracle.accumulate(res, [0, 0, 1, 1, 1], [0, 0, 1, 1, 1], t.names, t.resids, t.resnames,
                                  t.segids, s.backbone)
    assert np.array_equal(np.asarray(acc["key1"]), golden_synth[tag + "_acc_key1"])
    np.testing.assert_allclose(acc["score"], golden_synth[tag + "_acc_score"], rtol=1e-12, atol=1e-300)
