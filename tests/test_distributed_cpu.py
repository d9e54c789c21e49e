"""The N>1 path on CPU: world_size-2 (and 3) gloo groups.  Each rank "analyses" its frame slice
(rows built from the oracle stand in for the device output), then the ranks exchange results the way
bench.py / a multi-GPU run does: one all_reduce of the aggregated maps, one gather of the rows.
The merged result must equal the single-rank result -- the analogue of the reference's
test_multiCore_analysis (tests/test_basic.py:73-84)."""
import os
import socket

import numpy as np
import pytest

from oracle import frame_oracle
from pycontact_b200 import _lib, distributed, prepare
from pycontact_b200.results import AccumulationTable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rows_for_frames(frames_lo, frames_hi):
    """(frame, key) rows of the golden session's frames [lo, hi), frame index global."""
    from pycontact_b200.io.psf import Topology
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_golden.npz"))
    z = np.load(os.path.join(GOLDEN, "rpn11_ubq_input.npz"))
    names = [str(s) for s in z["names"]]
    resnames = [str(s) for s in z["resnames"]]
    segids = [str(s) for s in z["segids"]]
    resids = z["resids"].astype(np.int64)
    idx1, idx2 = g["sel1"].astype(np.int64), g["sel2"].astype(np.int64)
    gm1 = prepare.group_map(idx1, g["map1"], names, resids, resnames, segids)
    gm2 = prepare.group_map(idx2, g["map2"], names, resids, resnames, segids)
    loc1 = {int(a): k for k, a in enumerate(idx1)}
    loc2 = {int(a): k for k, a in enumerate(idx2)}
    bb = set(int(b) for b in z["backbone"])
    fp, hp = g["frame_ptr"], g["hb_ptr"]
    rows, frames = [], []
    for f in range(frames_lo, frames_hi):
        cells = {}
        for c in range(fp[f], fp[f + 1]):
            a, b = int(g["idx1"][c]), int(g["idx2"][c])
            key = int(gm1.gid[loc1[a]]) * gm2.n_groups + int(gm2.gid[loc2[b]])
            e = cells.setdefault(key, [0.0, 0.0, 0.0, 0, 0, c - fp[f]])
            w = float(g["weight"][c])
            e[0] += w
            e[1] += w if a in bb else 0.0
            e[2] += w if b in bb else 0.0
            e[3] += 1
            e[4] += int(hp[c + 1] - hp[c])
        for key, e in cells.items():
            rows.append((key, e[0], e[1], e[2], e[3], e[4], e[5]))
            frames.append(f)
    return (np.asarray(rows, _lib.row_dtype), np.asarray(frames, np.int64), gm1, gm2,
            (names, resids, resnames, segids), g)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = distributed.frame_slice(50, rank, world)
        rows, frames, gm1, gm2, topo, g = _rows_for_frames(lo, hi)
        n_keys = gm1.n_groups * gm2.n_groups
        totals = distributed.all_reduce_totals(distributed.aggregate_rows(rows, n_keys))
        all_rows, all_frames = distributed.gather_rows(rows, frames, dst=0)
        if rank == 0:
            table = AccumulationTable(50, gm1, gm2)
            table.add_rows(all_rows, all_frames)
            table.finalize()
            np.savez(out, score=table.score_matrix(), hbonds=table.hbond_matrix(), keys=table.keys,
                     totals=totals, hbp=table.hbond_percentage())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_frame_sharded_ranks_reproduce_single_rank(world, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    z = np.load(out)
    g = np.load(os.path.join(GOLDEN, "rpn11_ubq_golden.npz"))
    assert z["score"].shape == (148, 50)                                        # tests/test_basic.py:79
    np.testing.assert_allclose(z["score"], g["acc_score"], rtol=1e-12, atol=1e-300)   # same key ORDER too
    assert np.array_equal(z["hbonds"], g["acc_hbonds"])
    assert z["hbp"].sum() == 676.0                                              # tests/test_basic.py:84
    # the all-reduced aggregated maps equal the sums over the full score matrix
    tot = z["totals"].reshape(-1, 4)
    keys = z["keys"].astype(np.int64)
    np.testing.assert_allclose(tot[keys, 0], g["acc_score"].sum(axis=1), rtol=1e-12)
    np.testing.assert_allclose(tot[keys, 1], g["acc_bb1"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(tot[keys, 2], g["acc_bb2"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(tot[keys, 3], (g["acc_hbonds"] > 0).sum(axis=1))
    assert np.count_nonzero(tot[:, 0]) == 148


def test_frame_slices_cover_all_frames_once():
    for world in (1, 2, 3, 4, 8):
        got = []
        for r in range(world):
            lo, hi = distributed.frame_slice(50, r, world)
            got.extend(range(lo, hi))
        assert got == list(range(50))
    assert distributed.frame_slice(3, 7, 8) in ((0, 0), (2, 3), (1, 2))          # more ranks than frames
