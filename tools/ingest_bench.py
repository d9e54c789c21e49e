#!/usr/bin/env python
"""Ingest step in front of the hot path: DCD file -> selection gather -> pinned host memory
-> H2D -> scan.  Measures, on a synthetic two-chain system written as a CHARMM DCD file:

  read       DCDSource.read alone (memory-mapped file, gather of the two selections): frames/s, GB/s of selection
             coordinates produced and GB/s of file bytes touched
  scan       run_scan over the file without and with the reader thread (scheduler.prefetched): frames/s end to end
             (file -> contacts, H-bonds and accumulated rows on the host)

    python tools/ingest_bench.py [--system c2|small] [--frames 2000] [--dir /dev/shm]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycontact_b200 import prepare, synth
from pycontact_b200.io.dcd import DCDReader
from pycontact_b200.scheduler import DCDSource, run_scan


def write_dcd(path, system, nframes, chunk=64):
    n = system.topology.n_atoms
    with open(path, "wb") as fh:
        icntrl = np.zeros(20, "<i4"); icntrl[0] = nframes; icntrl[1] = 1; icntrl[2] = 1; icntrl[19] = 24
        fh.write(np.int32(84).tobytes() + b"CORD" + icntrl.tobytes() + np.int32(84).tobytes())
        title = b"synthetic trajectory of pycontact_b200".ljust(80)
        fh.write(np.int32(84).tobytes() + np.int32(1).tobytes() + title + np.int32(84).tobytes())
        fh.write(np.int32(4).tobytes() + np.int32(n).tobytes() + np.int32(4).tobytes())
        rec = np.int32(4 * n).tobytes()
        for f0 in range(0, nframes, chunk):
            co = synth.frames_host(system, range(f0, min(f0 + chunk, nframes)))
            for k in range(co.shape[0]):
                for ax in range(3):
                    fh.write(rec + np.ascontiguousarray(co[k, :, ax]).tobytes() + rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="c2")
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    ap.add_argument("--no-gpu", action="store_true")
    a = ap.parse_args()
    s = synth.make_system(a.system)
    t = s.topology
    path = os.path.join(a.dir, "pc_ingest_%s_%d.dcd" % (a.system, a.frames))
    t0 = time.perf_counter()
    write_dcd(path, s, a.frames)
    out = {"system": a.system, "atoms": t.n_atoms, "frames": a.frames, "file_GB": os.path.getsize(path) / 1e9,
           "write_s": time.perf_counter() - t0, "n1": len(s.sel1), "n2": len(s.sel2), "where": a.dir}
    src = DCDSource(DCDReader(path), s.sel1, s.sel2)
    bf = 256
    t0 = time.perf_counter()
    nb = 0
    for f0 in range(0, a.frames, bf):
        x1, x2 = src.read(f0, min(bf, a.frames - f0))
        nb += x1.nbytes + x2.nbytes
    dt = time.perf_counter() - t0
    out["read"] = {"frames_per_s": a.frames / dt, "selection_GBs": nb / dt / 1e9, "file_GBs": out["file_GB"] / dt,
                   "what": "DCDSource.read, batches of %d frames, one thread" % bf}
    if not a.no_gpu:
        from pycontact_b200.engine import ContactEngine
        hcsr = prepare.hydrogen_csr_from_bond_table(t.names, t.types, t.bonds)
        bb = np.zeros(t.n_atoms, bool); bb[s.backbone] = True
        seg = prepare.string_codes(t.segids)
        s1 = prepare.prepare_selection(s.sel1, t.names, t.resids, seg, bb, hcsr)
        s2 = prepare.prepare_selection(s.sel2, t.names, t.resids, seg, bb, hcsr)
        gm1 = prepare.group_map(s.sel1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
        gm2 = prepare.group_map(s.sel2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
        for pf in (0, 2):
            with ContactEngine(s1, s2, 5.0, 2.5, 120, slots=3) as eng:
                eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
                run_scan([eng], src, batch_frames=bf, order_keys=False, accumulate=True, group_maps=(gm1, gm2),
                         keep_contacts=True, prefetch=pf)                                   # warm-up (buffers, page cache)
                t0 = time.perf_counter()
                traj, tab, st = run_scan([eng], src, batch_frames=bf, order_keys=False, accumulate=True,
                                         group_maps=(gm1, gm2), keep_contacts=True, prefetch=pf)
                dt = time.perf_counter() - t0
            out["scan_prefetch_%d" % pf] = {"frames_per_s": a.frames / dt, "contacts": st["contacts"], "seconds": dt,
                                            "what": "run_scan file -> records on the host (incl. sorting into frame order), reader thread depth %d" % pf}
    os.remove(path)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
