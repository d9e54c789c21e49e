#!/usr/bin/env python
"""bench.py -- frames/s of the contact-analysis hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4s|c1] [--frames F]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU path on the host cores, same workload

A *step* is one pass of the hot path (neighbour search + filters + score + H-bond test +
per-frame accumulation + the end-of-run reduction of the aggregated maps) over one synthetic
trajectory of F frames per GPU (weak scaling: every rank analyses its own F frames).
``value``  : frames/s with the coordinates already in HBM (CUDA events on the launch stream).
``e2e``    : frames/s through the public Python API with HOST (pinned) coordinates: H2D copies,
             kernels and the D2H of all contacts / H-bonds / rows inside the timed region.
``roofline``: the dominant kernel's algorithmic bytes / its measured launch time, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.
``cpu_baseline``: the oracle port on the host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synthetic system, default frames per GPU per step, description)
    "c2": ("c2", 10000, "synthetic 100k-atom solvated protein (2 chains x 6000 atoms selected), "
                        "protein-protein interface contacts + H-bonds, cutoff 5.0 / 2.5 A / 120 deg"),
    "c3": ("c3", 712, "synthetic 1M-atom membrane-protein complex, protein (50k atoms) vs lipid+water (950k atoms)"),
    "c4s": ("c4_small", 256, "synthetic capsid-like shell, 64k atoms, all-vs-all self contacts"),
    "c4": ("c4", 16, "synthetic 3M-atom virus-capsid-like shell, all-vs-all residue contact map (self)"),
    "small": ("small", 2048, "small solvated two-chain system (debug)"),
}
METRIC = "frames/sec (contacts+H-bonds+accumulation)"


# ---------------------------------------------------------------------------------------------------
def prepare_workload(name):
    from pycontact_b200 import prepare, synth
    sysname, frames, desc = WORKLOADS[name]
    s = synth.make_system(sysname)
    t = s.topology
    self_mode = s.sel2text == "self"
    hcsr = prepare.hydrogen_csr_from_bond_table(t.names, t.types, t.bonds)
    bb = np.zeros(t.n_atoms, bool)
    bb[s.backbone] = True
    seg = prepare.string_codes(t.segids)
    s1 = prepare.prepare_selection(s.sel1, t.names, t.resids, seg, bb, hcsr)
    s2 = None if self_mode else prepare.prepare_selection(s.sel2, t.names, t.resids, seg, bb, hcsr)
    gm1 = prepare.group_map(s.sel1, s.maps[0], t.names, t.resids, t.resnames, t.segids)
    gm2 = gm1 if self_mode else prepare.group_map(s.sel2, s.maps[1], t.names, t.resids, t.resnames, t.segids)
    return dict(system=s, s1=s1, s2=s2, gm1=gm1, gm2=gm2, self_mode=self_mode, frames=frames, desc=desc)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.proc = None
        self.lines = []

    def run(self):
        if os.environ.get("PC_NO_CLOCKS"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for ln in self.proc.stdout:
                self.lines.append(ln.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port over a process pool (the reference's Pool-over-chunks pattern,
# core/multi_trajectory.py:423-451), timed on a bounded sample.
# ---------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(name):
    from oracle import frame_oracle, native
    w = prepare_workload(name)
    s = w["system"]
    t = s.topology
    _CPU.update(w=w, s=s, t=t, hcsr=frame_oracle.bonded_hydrogens(t.names, t.types, t.bonds),
                search=native.ref_find_contacts if native.have_reference() else native.oracle_find_contacts)


def _cpu_chunk(frames):
    from oracle import frame_oracle
    from pycontact_b200 import synth
    s, t, w = _CPU["s"], _CPU["t"], _CPU["w"]
    idx2 = s.sel1 if w["self_mode"] else s.sel2
    atoms = s.sel1 if w["self_mode"] else np.concatenate([s.sel1, s.sel2])
    coords = synth.frames_host(s, frames, atoms=atoms)
    n1 = len(s.sel1)
    t0 = time.perf_counter()
    res = []
    for k in range(len(frames)):
        x1 = coords[k, :n1]
        x2 = x1.copy() if w["self_mode"] else coords[k, n1:]
        res.append(frame_oracle.frame_contacts(x1, x2, s.sel1, idx2, 5.0, 2.5, 120, t.names, _CPU["hcsr"],
                                               t.resids, t.segids, w["self_mode"], _CPU["search"]))
    acc = frame_oracle.accumulate(res, s.maps[0], s.maps[1], t.names, t.resids, t.resnames, t.segids, s.backbone)
    dt = time.perf_counter() - t0
    return len(frames), sum(len(r["idx1"]) for r in res), sum(len(r["hb_donor"]) for r in res), len(acc["keys"]), dt


def cpu_run(name, frames_per_step, steps, warmup, cores):
    """Returns (frames/s over the timed steps, seconds per step, totals)."""
    import multiprocessing as mp
    from pycontact_b200.scheduler import chunks
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(name,)) as pool:
        def one_step(step):
            fr = list(range(step * frames_per_step, (step + 1) * frames_per_step))
            parts = [c for c in chunks(fr, cores) if len(c)]
            return pool.map(_cpu_chunk, parts)
        for wu in range(warmup):
            one_step(wu)
        t0 = time.perf_counter()
        out = [one_step(warmup + k) for k in range(steps)]
        dt = time.perf_counter() - t0
    contacts = sum(p[1] for st in out for p in st)
    hbonds = sum(p[2] for st in out for p in st)
    return steps * frames_per_step / dt, dt / steps, dict(contacts=contacts, hbonds=hbonds)


def sample_size(name, cores):
    """Frames per CPU step: ~1-4 s of pool work (the port costs ~0.1 s/frame on the 100k system)."""
    per_core = {"c2": 8, "c3": 1, "c4s": 1, "small": 32}[name]
    return per_core * cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = sample_size(args.workload, cores)
    from oracle import native
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    fps, sec, tot = cpu_run(args.workload, frames, steps, warmup, cores)
    sample = ("%d frames per step over a %d-process pool (chunks over frames like core/multi_trajectory.py:423-451); "
              "search = %s, pair filters / score / H-bond / accumulation = numpy restatement (oracle/frame_oracle.py)"
              % (frames, cores, "compiled reference gridsearch.C (oracle/_ref)" if native.have_reference()
                 else "C restatement (oracle/gridsearch_oracle.c)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 search / f64 score", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, WORKLOADS[args.workload][2]), "frames_per_step": frames},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "contacts_per_s": tot["contacts"] / (sec * steps), "hbonds_per_s": tot["hbonds"] / (sec * steps),
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu(args):
    import ctypes
    from pycontact_b200 import _lib, device_synth
    from pycontact_b200._lib import DeviceBuffer, check
    from pycontact_b200.engine import ContactEngine
    from pycontact_b200.scheduler import ArraySource, run_scan

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lib = _lib.load()
    if lib.pc_device_count() <= local:
        raise RuntimeError("bench.py needs a CUDA device per rank; there is no CPU fallback")
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = prepare_workload(args.workload)
    s, s1, s2, gm1, gm2 = w["system"], w["s1"], w["s2"], w["gm1"], w["gm2"]
    self_mode = w["self_mode"]
    F = args.frames or w["frames"]
    n1, n2 = s1.n, (s1.n if self_mode else s2.n)
    frame0 = rank * F                                      # every rank analyses its own F frames
    gen1 = device_synth.DeviceTrajectory(s, local, s.sel1)
    gen2 = None if self_mode else device_synth.DeviceTrajectory(s, local, s.sel2)
    d1 = DeviceBuffer(F * n1 * 12, local)
    d2 = None if self_mode else DeviceBuffer(F * n2 * 12, local)
    gen1.generate(d1.ptr, frame0, F)
    if gen2 is not None:
        gen2.generate(d2.ptr, frame0, F)
    check(lib.pc_device_sync(local))

    eng = ContactEngine(s1, s2, 5.0, 2.5, 120, device=local, slots=args.slots)
    eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
    n_keys_all = gm1.n_groups * gm2.n_groups
    if n_keys_all <= (1 << 20) or (args.sparse_fold and n_keys_all < (1 << 32)):
        # bb sums travel in the time-aggregated table (dense or sparse) instead of per (frame, key) cell
        eng.set_pack_format(12)
    in_bytes_frame = (n1 + (0 if self_mode else n2)) * 12
    batch = args.batch or int(max(1, min(F, 2048, (2 << 30) // in_bytes_frame)))
    nb = (F + batch - 1) // batch
    n_keys = gm1.n_groups * gm2.n_groups
    dense = n_keys <= (1 << 20)          # aggregated maps as a dense [n_keys][4] table (32 MB at most)
    sparse = args.sparse_fold and (not dense) and n_keys < (1 << 32)   # opt-in: a hash table per slot, merged per step
    sparse_entries = args.sparse_entries
    # `value` keeps all per-frame records (contacts, H-bonds, rows) in HBM, like its inputs; what goes to the
    # host is the time-aggregated map (after the all-reduce over ranks).  Without such a map (key spaces too
    # large for the dense table and no --sparse-fold) the packed per-frame rows are copied to the host instead.
    # `e2e` always returns every record to the host.
    rows_to_host = args.rows_to_host or not (dense or sparse)
    totals = DeviceBuffer(max(1, n_keys if dense else 1) * 32, local)
    totals_t = None
    if world > 1 and dense:
        totals_t = torch.zeros(n_keys * 4, dtype=torch.float64, device="cuda")
    totals_host = _lib.PinnedArray((max(1, n_keys if dense else 1) * 4,), np.float64)
    stream = eng.slots[0].stream
    ev = [[ctypes.c_void_p(), ctypes.c_void_p()] for _ in range(nb + 1)]
    for pair in ev:
        for e in pair:
            check(lib.pc_event_create(local, ctypes.byref(e)))

    def totals_ptr():
        return ctypes.c_void_p(totals_t.data_ptr()) if totals_t is not None else totals.ptr

    def device_step(timed):
        """One pass over the F resident frames.  Returns per-step counts."""
        tot = np.zeros(4, np.int64)
        d_rows = 0
        if dense:
            if totals_t is not None:
                totals_t.zero_()
                torch.cuda.current_stream().synchronize()
            else:
                check(lib.pc_memset(local, totals.ptr, 0, n_keys * 32, stream))
        if sparse:
            eng.fold_sparse_init(sparse_entries)
        trace = bool(os.environ.get("PC_TRACE"))

        def finish(sid, nf):
            nonlocal d_rows
            _t0 = time.perf_counter()
            c = eng.collect_device(sid)
            _t1 = time.perf_counter()
            tot[:] += c
            nrows = 0
            if rows_to_host:
                rows, _, _ = eng.fetch_rows(sid, packed=True)  # the per-frame maps of this rank's frames -> host
                d_rows += rows.nbytes
                nrows = rows.size
            if trace:
                print("slot %d nf %d: wait %.3f ms, fold+fetch %.3f ms (%d rows)" % (
                    sid, nf, (_t1 - _t0) * 1e3, (time.perf_counter() - _t1) * 1e3, nrows), file=sys.stderr)

        # several batches in flight (one per slot): later batches are scanned while an earlier batch's
        # rows travel to the host
        pending = []
        for b in range(nb):
            f0 = b * batch
            nf = min(batch, F - f0)
            sid = b % len(eng.slots)
            if len(pending) == len(eng.slots):
                finish(*pending.pop(0))
            _ts = time.perf_counter()
            eng.submit_device(sid, d1.at(f0 * n1 * 12), None if self_mode else d2.at(f0 * n2 * 12), nf,
                              accumulate=True, events=ev[b] if timed else None, pack_rows=True,
                              fold=(totals_ptr(), n_keys) if dense else ("sparse" if sparse else None))
            if trace:
                print("submit slot %d: %.3f ms" % (sid, (time.perf_counter() - _ts) * 1e3), file=sys.stderr)
            pending.append((sid, nf))
        while pending:
            finish(*pending.pop(0))
        if dense:
            for sl in eng.slots:
                check(lib.pc_stream_sync(sl.stream))
            if totals_t is not None:
                dist.all_reduce(totals_t)                  # the one collective: aggregated maps over ranks
                torch.cuda.current_stream().synchronize()
            if rank == 0:
                host = totals_host.array
                src = totals_ptr()
                check(lib.pc_memcpy(local, host.ctypes.data, src, host.nbytes, 2, None))
                check(lib.pc_device_sync(local))
                d_rows += host.nbytes
        if sparse:
            _tf = time.perf_counter()
            ent = eng.fold_sparse_fetch()                  # per-key sums of this rank's frames
            d_rows += ent.nbytes
            if trace:
                print("sparse fetch+merge %.3f ms, %d keys" % ((time.perf_counter() - _tf) * 1e3, ent.size), file=sys.stderr)
            if dist is not None:                           # ranks hold disjoint frames: sum per key on rank 0
                cnt = torch.tensor([ent.size], device="cuda")
                dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
                pad = np.zeros(int(cnt.item()), ent.dtype)
                pad["key"] = np.uint64(0xffffffffffffffff)
                pad[:ent.size] = ent
                mine = torch.from_numpy(pad.view(np.uint8).copy()).cuda()
                allv = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
                dist.gather(mine, allv, dst=0)
                if rank == 0:
                    merged = np.concatenate([t.cpu().numpy().view(ent.dtype) for t in allv])
                    merged = merged[merged["key"] != np.uint64(0xffffffffffffffff)]
                    keys, inv = np.unique(merged["key"], return_inverse=True)
                    ent = np.zeros(keys.size, ent.dtype)
                    ent["key"] = keys
                    for fld in ("score", "bb1", "bb2", "hbframes"):
                        ent[fld] = np.bincount(inv, weights=merged[fld], minlength=keys.size)
            device_step.last_entries = int(ent.size)
        return tot, d_rows


    def barrier():
        if dist is not None:
            dist.barrier()
        check(lib.pc_device_sync(local))

    for _ in range(max(args.warmup, 3 if not args.quick else 1)):
        device_step(False)
    launches0 = lib.pc_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    barrier()
    e_all = ev[nb]
    check(lib.pc_event_record(e_all[0], stream))
    t0 = time.perf_counter()
    kernel_ms = 0.0
    tot = np.zeros(4, np.int64)
    for k in range(args.steps):
        c, d_bytes = device_step(False)
        tot += c
    check(lib.pc_event_record(e_all[1], stream))
    ms = ctypes.c_float()
    check(lib.pc_event_elapsed_ms(e_all[0], e_all[1], ctypes.byref(ms)))
    # The dominant kernel alone: the same launches again, one at a time on one stream (in the step above
    # two batches are in flight, so an event pair around one scan also sees the other stream's kernels).
    kernel_ms = 0.0
    k_launches = 0
    large = eng_path(eng) == 2
    probe_ms = 0.0
    pe = [ctypes.c_void_p(), ctypes.c_void_p()]
    if large:
        for e in pe:
            check(lib.pc_event_create(local, ctypes.byref(e)))
        check(lib.pc_set_probe(eng.slots[0].ctx, pe[0], pe[1]))
    for k in range(max(1, min(args.steps, 3))):
        for b in range(nb):
            f0b = b * batch
            nfb = min(batch, F - f0b)
            eng.submit_device(0, d1.at(f0b * n1 * 12), None if self_mode else d2.at(f0b * n2 * 12), nfb,
                              accumulate=False, events=ev[b])
            eng.collect_device(0)
            kms = ctypes.c_float()
            check(lib.pc_event_elapsed_ms(ev[b][0], ev[b][1], ctypes.byref(kms)))
            kernel_ms += kms.value
            k_launches += 1
            if large:
                check(lib.pc_event_elapsed_ms(pe[0], pe[1], ctypes.byref(kms)))
                probe_ms += kms.value
    if large:
        check(lib.pc_set_probe(eng.slots[0].ctx, None, None))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = lib.pc_launch_count() - launches0
    dev_s = max(ms.value / 1e3, 1e-9)
    # the step ends with host-visible results, so the device time is bounded by the wall time of
    # the same region; report the device-event time, max over ranks
    if dist is not None:
        tt = torch.tensor([dev_s, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, wall = float(tt[0]), float(tt[1])
        cc = torch.tensor(tot, device="cuda")
        dist.all_reduce(cc)
        tot_all = cc.cpu().numpy()
    else:
        tot_all = tot
    frames_total = F * world * args.steps
    value = frames_total / dev_s
    contacts_step, hbonds_step, rows_step, evals_step = (tot / args.steps).tolist()

    # ---- roofline of the dominant kernel (the scan) -------------------------------------------------
    peaks, peak_kind = measured_peaks()
    alg_bytes_step = F * in_bytes_frame + 16 * contacts_step + 20 * hbonds_step      # read coords once, write records
    k_s = kernel_ms / 1e3 / k_launches
    scan_gbs = alg_bytes_step / nb / k_s / 1e9          # all kernels of one scan launch sequence
    if large:
        # dominant kernel of the large-frame path: the streaming bin over the bigger selection, which reads
        # that selection's packed coordinates once (12 B/atom) -- its own algorithmic bytes and duration
        n_dom = n1 if self_mode else n2
        k_bytes = 12.0 * n_dom * F / nb
        k_s_dom = probe_ms / 1e3 / k_launches
        achieved = k_bytes / k_s_dom / 1e9
    else:
        k_bytes = alg_bytes_step / nb
        k_s_dom = k_s
        achieved = scan_gbs
    sm_ghz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) / 1e3
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh).get(args.workload)
        if tj:
            traffic = tj["dram_bytes_per_frame"] * (F / nb)
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                "kernel": "pc_scan_small_kernel" if not large else "pc_large_stream_kernel<BIN> (streaming bin of the larger selection)",
                "kernel_share_of_step": (k_s_dom * nb) / (dev_s / args.steps),
                "kernel_ms_per_launch": k_s_dom * 1e3,
                "algorithmic_bytes_per_launch": k_bytes,
                "scan_sequence_ms_per_batch": k_s * 1e3, "scan_sequence_gbs": scan_gbs,
                "step_gbs": (alg_bytes_step + 16 * rows_step) / (dev_s / args.steps) / 1e9,
                "pair_evals_per_s": evals_step / nb / k_s,
                "fp32_pair_eval_peak_per_s": 148 * 128 * sm_ghz * 1e9 / 9.0}

    # ---- e2e: the public API with host buffers (rank-local frames) ------------------------------------
    Fe = min(F, args.e2e_frames or F)
    h1 = _lib.PinnedArray((Fe, n1, 3), np.float32)
    check(lib.pc_memcpy(local, h1.array.ctypes.data, d1.ptr, Fe * n1 * 12, 2, None))
    h2 = None
    if not self_mode:
        h2 = _lib.PinnedArray((Fe, n2, 3), np.float32)
        check(lib.pc_memcpy(local, h2.array.ctypes.data, d2.ptr, Fe * n2 * 12, 2, None))
    check(lib.pc_device_sync(local))
    src = ArraySource(h1.array, None if h2 is None else h2.array)
    eb = int(max(1, min(Fe, (64 << 20) // in_bytes_frame)))

    def e2e_step():
        gen = eng.scan_batches(((f0, *src.read(f0, min(eb, Fe - f0))) for f0 in range(0, Fe, eb)),
                               accumulate=True, pinned=True, copy=False)
        d2h = 0
        for r in gen:
            d2h += r.contacts.nbytes + r.hbonds.nbytes + r.rows.nbytes + 6 * 4 * r.nframes
        return d2h
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_steps = max(1, min(args.steps, 5))
    for _ in range(e_steps):
        d2h = e2e_step()
    barrier()
    e_s = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e_s = float(tt[0])
    e2e = {"value": Fe * world * e_steps / e_s, "unit": "frames/s", "h2d_bytes_per_step": int(Fe * in_bytes_frame),
           "d2h_bytes_per_step": int(d2h), "frames_per_step": Fe, "steps": e_steps,
           "api": "ContactEngine.scan_batches (pinned host frames -> contacts, H-bonds, rows on the host)"}

    # ---- CPU baseline on rank 0 (bounded sample) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        from oracle import native
        fr = sample_size(args.workload, cores)
        fps, sec, _ = cpu_run(args.workload, fr, 2, 1, cores)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "2 steps x %d frames of the same synthetic workload over a %d-process pool; search = %s, "
                         "rest = numpy restatement of the reference's pair loop (oracle/frame_oracle.py)"
                         % (fr, cores, "compiled reference gridsearch.C (oracle/_ref)" if native.have_reference()
                            else "oracle/gridsearch_oracle.c")}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3 if not args.quick else 1), "ms_per_step": dev_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (pair test, exact order) + f64 (distance, weight, H-bond geometry, sums)",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, w["desc"]), "frames_per_gpu_per_step": F,
                       "atoms_sel1": n1, "atoms_sel2": n2, "batch_frames": batch, "batches_per_step": nb,
                       "l2": "inputs (%.0f MB per step) exceed the 126 MB L2" % (F * in_bytes_frame / 1e6),
                       "results": ("per-frame records stay in HBM; " if not rows_to_host else "packed per-frame rows -> host; ")
                                  + ("aggregated maps (dense table) all-reduced and copied to the host" if dense else
                                     "aggregated maps (sparse table) merged on the host" if sparse else "no aggregated table"),
                       "parallelism": "frames sharded over %d GPU(s), one all-reduce of the aggregated maps" % world},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "per_step": {"contacts": contacts_step, "hbonds": hbonds_step, "rows": rows_step,
                         "pair_evals": evals_step, "wall_ms": wall / args.steps * 1e3},
            "contacts_per_s": float(tot_all[0]) / dev_s, "hbonds_per_s": float(tot_all[1]) / dev_s,
            "pair_evals_per_s": float(tot_all[3]) / dev_s,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def eng_path(eng):
    import ctypes
    from pycontact_b200._lib import check
    p = ctypes.c_int32()
    check(eng.lib.pc_last_status(eng.slots[0].ctx, None, None, None, None, ctypes.byref(p)))
    return int(p.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: per workload)")
    ap.add_argument("--batch", type=int, default=0, help="frames per launch")
    ap.add_argument("--e2e-frames", type=int, default=0)
    ap.add_argument("--slots", type=int, default=3, help="frame batches in flight per GPU")
    ap.add_argument("--sparse-entries", type=int, default=1 << 21)
    ap.add_argument("--rows-to-host", action="store_true",
                    help="value: also copy the packed per-frame rows to the host inside the timed step")
    ap.add_argument("--sparse-fold", action="store_true",
                    help="large key spaces: 12-byte rows + on-device sparse aggregation (less D2H, more device work)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--quick", action="store_true", help="profiling runs: 1 warm-up step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
