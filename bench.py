#!/usr/bin/env python
"""bench.py -- frames/s of the contact-analysis hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c2i|c4|c4s|c1|small] [--frames F]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU path on the host cores, same workload

Default workload: ``c3`` -- the 1 M-atom membrane-protein configuration BASELINE.json's target names (protein vs
lipid + water, 10 000 frames = 120 GB of float32 coordinates resident in HBM on one GPU; ``--gpus N`` cuts the SAME
trajectory into N contiguous slices like the reference's ``chunks`` over pool workers: strong scaling).

A *step* is one pass of the hot path (neighbour search + heavy-atom / self filters + distance + score + H-bond test
+ per-frame accumulation + the time-aggregated maps, reduced over ranks) over the whole trajectory.
``value``   : frames/s with the coordinates already in HBM (CUDA events on the launch stream, max over ranks).
``e2e``     : frames/s through the public Python API with HOST (pinned) coordinates: H2D copies, kernels and the
              D2H of all contacts / H-bonds / rows inside the timed region.
``roofline``: ``frac`` = the WHOLE step's algorithmic bytes (12 B per selected atom and frame + 16 B per contact +
              20 B per H-bond + 16 B per (frame, key) cell) / step time / measured HBM copy bandwidth
              (MEASURED_PEAKS.json); the dominant kernel's own figures are under ``dominant_kernel``, the FP32
              pair-evaluation roof (measured FFMA rate / 9 instructions) under ``fp32``.
``checked`` : per-frame contacts / H-bonds / sum of weights of the frames the CPU leg analysed, GPU vs CPU.
``cpu_baseline``: the UNMODIFIED reference loop (baseline/_ref) on a bounded sample; the numpy port beside it.
"""
from __future__ import annotations

import argparse
import ctypes
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
    # name: (system, frames of the whole trajectory, description)
    "c1": ("fixture", 3200, "rpn11_ubq fixture (4327 atoms, Rpn11 3097 atoms vs ubiquitin 1230 atoms), its 50 frames "
                            "repeated 64 times, 5.0 A cutoff + H-bond test (tests/test_basic.py:32-43)"),
    "c2": ("c2", 10000, "synthetic 100k-atom solvated protein (2 chains x 6000 atoms selected), "
                        "protein-protein interface contacts + H-bonds, cutoff 5.0 / 2.5 A / 120 deg"),
    "c2i": ("c2i", 10000, "the c2 system with the two chains interdigitated (residues alternate on one lattice: the "
                          "selections' bounding boxes coincide, every atom is inside the search grid)"),
    "c3": ("c3", 10000, "synthetic 1M-atom membrane-protein complex, protein (50k atoms) vs lipid+water (950k atoms), "
                        "10k frames, cutoff 5.0 / 2.5 A / 120 deg, maps resid+resname"),
    "c4s": ("c4_small", 256, "synthetic capsid-like shell, 64k atoms, all-vs-all self contacts"),
    "c4": ("c4", 64, "synthetic 3M-atom virus-capsid-like shell, all-vs-all residue contact map (self)"),
    "small": ("small", 2048, "small solvated two-chain system (debug)"),
}
METRIC = "frames/sec (contacts+H-bonds+accumulation)"
DTYPE = "f32 (pair test, exact order) + f64 (distance, weight, H-bond geometry)"
CUTOFF, HB_CUTOFF, HB_ANGLE = 5.0, 2.5, 120


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
class _FixtureSystem:
    """BASELINE.json configs[0]: the reference's rpn11_ubq example as committed arrays (tests/golden)."""

    def __init__(self):
        from pycontact_b200.io.psf import Topology
        from pycontact_b200.io.selection import select_atoms
        z = np.load(os.path.join(ROOT, "tests", "golden", "rpn11_ubq_input.npz"))
        self.name, self.seed = "fixture", 0
        self.coords = z["coords"]
        n = len(z["names"])
        self.topology = Topology([str(s) for s in z["segids"]], z["resids"].astype(np.int64), [str(s) for s in z["resnames"]],
                                 [str(s) for s in z["names"]], [str(s) for s in z["types"]], np.zeros(n), np.ones(n),
                                 z["bonds"].astype(np.int64))
        self.backbone = z["backbone"].astype(np.int64)
        self.sel1text, self.sel2text = "segid RN11", "segid UBQ"
        self.sel1 = select_atoms(self.topology, self.sel1text)
        self.sel2 = select_atoms(self.topology, self.sel2text)
        self.maps = ([0, 0, 1, 1, 0], [0, 0, 1, 1, 0])
        self.n_frames = self.coords.shape[0]

    def frames(self, frames, atoms):
        return np.ascontiguousarray(self.coords[np.asarray(frames) % self.n_frames][:, atoms])


class _FixtureGen:
    def __init__(self, system, device, atoms):
        self.system, self.device, self.atoms, self.n = system, device, np.asarray(atoms), len(atoms)

    def generate(self, out_ptr, frame0, nframes):
        from pycontact_b200 import _lib
        lib = _lib.load()
        for lo in range(0, nframes, 512):
            k = min(512, nframes - lo)
            h = self.system.frames(range(frame0 + lo, frame0 + lo + k), self.atoms)
            _lib.check(lib.pc_memcpy(self.device, ctypes.c_void_p(out_ptr.value + lo * self.n * 12), h.ctypes.data, h.nbytes, 1, None))
            _lib.check(lib.pc_device_sync(self.device))


def prepare_workload(name):
    from pycontact_b200 import prepare, synth
    sysname, frames, desc = WORKLOADS[name]
    s = _FixtureSystem() if sysname == "fixture" else synth.make_system(sysname)
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
    return dict(name=name, system=s, s1=s1, s2=s2, gm1=gm1, gm2=gm2, self_mode=self_mode, frames=frames, desc=desc)


def host_frames(w, frames):
    """Selection coordinates of trajectory frames from the host generator (the CPU arm's input)."""
    from pycontact_b200 import synth
    s = w["system"]
    if isinstance(s, _FixtureSystem):
        return s.frames(frames, s.sel1), (None if w["self_mode"] else s.frames(frames, s.sel2))
    atoms = s.sel1 if w["self_mode"] else np.concatenate([s.sel1, s.sel2])
    c = synth.frames_host(s, frames, atoms=atoms)
    n1 = len(s.sel1)
    return np.ascontiguousarray(c[:, :n1]), (None if w["self_mode"] else np.ascontiguousarray(c[:, n1:]))


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
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# CPU legs.  (1) The UNMODIFIED reference loop from baseline/_ref (baseline/ref_runner.py); (2) the oracle
# port over a process pool.  Both on a bounded sample of the same workload.
# ---------------------------------------------------------------------------------------------------
def frame_summaries_oracle(w, x1, x2, hcsr=None, search=None):
    """[(contacts, hbonds, sum of weights)] per frame from the numpy restatement (oracle/frame_oracle.py)."""
    from oracle import frame_oracle, native
    s = w["system"]
    t = s.topology
    hcsr = hcsr or frame_oracle.bonded_hydrogens(t.names, t.types, t.bonds)
    search = search or (native.ref_find_contacts if native.have_reference() else native.oracle_find_contacts)
    idx2 = s.sel1 if w["self_mode"] else s.sel2
    if max(len(s.sel1), len(idx2)) > 1500000:
        # 3 M atoms: the raw pair list with hydrogens (10^8 pairs) does not fit; the contact set is the same
        is_h = np.asarray([nm.startswith("H") for nm in t.names])
        search = frame_oracle.heavy_only_search(is_h[s.sel1], is_h[idx2], search)
    out = []
    for k in range(len(x1)):
        b = x1[k].copy() if w["self_mode"] else x2[k]
        r = frame_oracle.frame_contacts(x1[k], b, s.sel1, idx2, CUTOFF, HB_CUTOFF, HB_ANGLE, t.names, hcsr, t.resids, t.segids,
                                        w["self_mode"], search)
        out.append((int(len(r["idx1"])), int(len(r["hb_donor"])), float(r["weight"].sum())))
    return out


_CPU = {}


def _cpu_init(name):
    from oracle import frame_oracle
    w = prepare_workload(name)
    t = w["system"].topology
    _CPU.update(w=w, hcsr=frame_oracle.bonded_hydrogens(t.names, t.types, t.bonds))


def _cpu_chunk(frames):
    from oracle import frame_oracle
    w = _CPU["w"]
    s = w["system"]
    t = s.topology
    x1, x2 = host_frames(w, frames)
    t0 = time.perf_counter()
    from oracle import native
    search = native.ref_find_contacts if native.have_reference() else native.oracle_find_contacts
    idx2 = s.sel1 if w["self_mode"] else s.sel2
    res = []
    for k in range(len(frames)):
        b = x1[k].copy() if w["self_mode"] else x2[k]
        res.append(frame_oracle.frame_contacts(x1[k], b, s.sel1, idx2, CUTOFF, HB_CUTOFF, HB_ANGLE, t.names, _CPU["hcsr"],
                                               t.resids, t.segids, w["self_mode"], search))
    c1 = frame_oracle.atom_key_codes(s.maps[0], t.n_atoms, t.names, t.resids, t.resnames, t.segids)
    c2 = c1 if s.maps[1] == s.maps[0] else frame_oracle.atom_key_codes(s.maps[1], t.n_atoms, t.names, t.resids, t.resnames, t.segids)
    frame_oracle.accumulate_arrays(res, c1, c2, s.backbone, t.n_atoms)
    dt = time.perf_counter() - t0
    return len(frames), sum(len(r["idx1"]) for r in res), sum(len(r["hb_donor"]) for r in res), dt


def port_pool_run(name, frames_per_step, steps, warmup, cores):
    """The oracle port over a process pool (the reference's Pool-over-chunks pattern,
    core/multi_trajectory.py:423-451).  Returns (frames/s, seconds per step, totals)."""
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


# frames per leg, sized for ~10-30 s of CPU work each (per-frame cost of the reference loop measured in the build
# container: c1 0.14 s, c2 ~0.4 s, c3 6.6 s; every pool task of the reference pickles the whole bond table --
# ~20 s per task at 1 M atoms)
REF_SAMPLE = {"c1": (16, 16, 0), "c2": (8, 16, 0), "c2i": (8, 16, 0), "c3": (2, 2, 2), "c4s": (2, 4, 0), "c4": (0, 0, 0),
              "small": (16, 16, 0)}      # (frames at nproc = 1, frames through the pool, pool processes; 0 = cpu_count)
PORT_PER_CORE = {"c1": 8, "c2": 8, "c2i": 8, "c3": 1, "c4s": 1, "c4": 0, "small": 32}


def reference_loop_legs(w, x1, x2, cores, with_pool=True):
    """The unmodified reference on frames x1/x2: serial, through its LoggingPool, search only.  Returns a dict of
    figures and the per-frame summaries of the serial leg."""
    from baseline import ref_runner
    s = w["system"]
    t = s.topology
    n_serial, n_pool, pool_procs = REF_SAMPLE[w["name"]]
    pool_procs = pool_procs or cores
    t0 = time.perf_counter()
    bonds = ref_runner.bond_table(t)
    setup_s = time.perf_counter() - t0
    idx2 = s.sel1 if w["self_mode"] else s.sel2
    f1 = [x1[k] for k in range(len(x1))]
    f2 = [x1[k].copy() if w["self_mode"] else x2[k] for k in range(len(x1))]
    out = {"bond_table_s": setup_s}
    ns = min(n_serial, len(f1))
    res, dt = ref_runner.run_chunks(f1[:ns], f2[:ns], s.sel1, idx2, t, bonds, w["self_mode"], 1, CUTOFF, HB_CUTOFF, HB_ANGLE)
    out["nproc1"] = {"frames_per_s": ns / dt, "frames": ns, "seconds": dt}
    summaries = [ref_runner.frame_summary(fr) for fr in res]
    sdt, npairs = ref_runner.search_only(f1[:ns], f2[:ns], CUTOFF)
    out["search_only_nproc1"] = {"frames_per_s": ns / sdt, "raw_pairs_per_frame": npairs / max(ns, 1)}
    if with_pool and n_pool:
        npf = min(n_pool, len(f1))
        _, pdt = ref_runner.run_chunks(f1[:npf], f2[:npf], s.sel1, idx2, t, bonds, w["self_mode"], pool_procs, CUTOFF, HB_CUTOFF,
                                       HB_ANGLE)
        out["pool"] = {"frames_per_s": npf / pdt, "frames": npf, "seconds": pdt, "nproc": pool_procs}
    return out, summaries


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import native
    cores = os.cpu_count() or 1
    if native.have_reference():
        native.ref_find_contacts(np.zeros((1, 3), np.float32), np.ones((1, 3), np.float32), 5.0)   # dlopen in THIS process
    w = prepare_workload(args.workload)
    name = args.workload
    from baseline import ref_runner
    line = {"impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 search / f64 score", "data": "synthetic",
            "config": {"workload": "%s: %s" % (name, w["desc"])}}
    legs = None
    steps = 1
    if ref_runner.available() and REF_SAMPLE[name][0]:
        n_serial, n_pool, _ = REF_SAMPLE[name]
        x1, x2 = host_frames(w, list(range(max(n_serial, n_pool))))
        t0 = time.perf_counter()
        legs, _ = reference_loop_legs(w, x1, x2, cores)
        best = max([legs["nproc1"]] + ([legs["pool"]] if "pool" in legs else []), key=lambda d: d["frames_per_s"])
        fps = best["frames_per_s"]
        used = best.get("nproc", 1)
        sample = ("UNMODIFIED reference (baseline/_ref: loop_trajectory_grid + compiled cy_find_contacts, "
                  "PyContact/core/multi_trajectory.py:50-222) on frames 0..%d of the same synthetic trajectory: nproc=1 on %d "
                  "frames, LoggingPool + chunks (:423-451) with %s processes on %d frames%s; value = the faster of the two"
                  % (max(n_serial, n_pool) - 1, legs["nproc1"]["frames"], legs.get("pool", {}).get("nproc", "-"),
                     legs.get("pool", {}).get("frames", 0),
                     " (every pool task pickles the 1M-entry bond table, ~20 s per task: a cpu_count-wide pool would take minutes per step and be slower still)"
                     if name == "c3" else ""))
        kind = "reference"
        ms = best["seconds"] * 1e3
        line["config"]["frames_per_step"] = best["frames"]
    else:
        fr = max(1, PORT_PER_CORE[name] * cores)
        steps = min(args.steps, 3)
        fps, sec, _ = port_pool_run(name, fr, steps, min(args.warmup, 1), cores)
        used, kind, ms = cores, "port", sec * 1e3
        sample = ("baseline/_ref absent: oracle port (oracle/frame_oracle.py around %s), %d frames per step over a %d-process pool"
                  % ("the compiled reference gridsearch.C" if native.have_reference() else "oracle/gridsearch_oracle.c", fr, cores))
        line["config"]["frames_per_step"] = fr
    port = None
    if PORT_PER_CORE[name] and not args.no_port:
        fr = PORT_PER_CORE[name] * cores
        pfps, psec, _ = port_pool_run(name, fr, 1, 0, cores)
        port = {"frames_per_s": pfps, "frames": fr, "nproc": cores,
                "what": "numpy restatement of the pair loop (oracle/frame_oracle.py) around the compiled reference search"}
    line.update({"value": fps, "steps": steps, "warmup": 0, "ms_per_step": ms,
                 "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": used, "kind": kind, "sample": sample,
                                  "host_cores": cores, "reference_loop": legs, "port": port},
                 "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# algorithmic pair evaluations: the reference grid restricted to heavy atoms
# ---------------------------------------------------------------------------------------------------
def algorithmic_pair_evals(x1, x2, heavy1, heavy2, cutoff):
    """E_alg of one frame: sum over boxes of nA_heavy(box) * sum over the 27 neighbour boxes of nB_heavy, on the
    reference's grid (origin = joint minimum over ALL atoms of both selections, box edge = (float)cutoff grown x1.26
    while there are more than 4e6 boxes, src/gridsearch.C:940-985; visiting pattern :1111-1142)."""
    allx = np.concatenate([x1, x2]) if x2 is not None else x1
    lo = allx.min(0).astype(np.float32)
    rng = (allx.max(0).astype(np.float32) - lo).astype(np.float32)
    pd = np.float32(cutoff)
    while True:
        inv = np.float32(1.0) / pd
        nb = (rng * inv).astype(np.int64) + 1
        if nb.prod() <= 4000000:
            break
        pd = np.float32(pd * np.float32(1.26))
    def hist(x, heavy):
        c = np.minimum(((x[heavy] - lo) * inv).astype(np.int64), nb - 1)
        return np.bincount((c[:, 2] * nb[1] + c[:, 1]) * nb[0] + c[:, 0], minlength=int(nb.prod())).reshape(nb[2], nb[1], nb[0])
    ha = hist(x1, heavy1).astype(np.float64)
    hb = ha if x2 is None else hist(x2, heavy2).astype(np.float64)
    pad = np.pad(hb, 1)
    nbr = np.zeros_like(hb)
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                nbr += pad[dz:dz + nb[2], dy:dy + nb[1], dx:dx + nb[0]]
    return float((ha * nbr).sum())


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu(args):
    from pycontact_b200 import _lib, device_synth
    from pycontact_b200._lib import DeviceBuffer, check
    from pycontact_b200.engine import ContactEngine
    from pycontact_b200.scheduler import ArraySource, worker_slices

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lib = _lib.load()
    if lib.pc_device_count() <= local:
        raise RuntimeError("bench.py needs a CUDA device per rank; there is no CPU fallback")
    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = prepare_workload(args.workload)
    s, s1, s2, gm1, gm2 = w["system"], w["s1"], w["s2"], w["gm1"], w["gm2"]
    self_mode = w["self_mode"]
    n1, n2 = s1.n, (s1.n if self_mode else s2.n)
    in_bytes_frame = (n1 + (0 if self_mode else n2)) * 12
    F_total = args.frames or w["frames"]
    # ---- which frames of the trajectory this rank analyses --------------------------------------------------
    if args.scaling == "strong":
        # one trajectory, cut like core/multi_trajectory.py:423-426 cuts it over pool workers
        lo, hi = worker_slices(F_total, world)[rank]
    else:
        lo, hi = rank * F_total, (rank + 1) * F_total          # weak: every rank its own F_total frames
    fr_b, to_b = ctypes.c_int64(), ctypes.c_int64()
    check(lib.pc_mem_info(local, ctypes.byref(fr_b), ctypes.byref(to_b)))
    fit = int((fr_b.value - (20 << 30)) // in_bytes_frame)      # leave 20 GB for work buffers and outputs
    clipped = None
    if hi - lo > fit:
        clipped = "%d frames per GPU do not fit %d GB of free HBM: %d analysed" % (hi - lo, fr_b.value >> 30, max(fit, 1))
        hi = lo + max(fit, 1)
    F = hi - lo
    if isinstance(s, _FixtureSystem):
        gen1 = _FixtureGen(s, local, s.sel1)
        gen2 = None if self_mode else _FixtureGen(s, local, s.sel2)
    else:
        gen1 = device_synth.DeviceTrajectory(s, local, s.sel1)
        gen2 = None if self_mode else device_synth.DeviceTrajectory(s, local, s.sel2)
    d1 = DeviceBuffer(F * n1 * 12, local)
    d2 = None if self_mode else DeviceBuffer(F * n2 * 12, local)
    for f0 in range(0, F, 1024):                                # the generator takes int32 element counts per launch
        nf = min(1024, F - f0)
        gen1.generate(d1.at(f0 * n1 * 12), lo + f0, nf)
        if gen2 is not None:
            gen2.generate(d2.at(f0 * n2 * 12), lo + f0, nf)
    check(lib.pc_device_sync(local))

    eng = ContactEngine(s1, s2, CUTOFF, HB_CUTOFF, HB_ANGLE, device=local, slots=args.slots)
    eng.set_maps(gm1.gid, gm2.gid, gm2.n_groups)
    if os.environ.get("PC_GROUP"):                         # experiments: 0 = multi-kernel sequence + global table instead of the group kernel
        eng.set_group_mode(int(os.environ["PC_GROUP"]))
    if os.environ.get("PC_SMALL_CTAS"):                    # experiments: CTA plan of the small-frame kernel (pc_set_small_ctas)
        for sl in eng.slots:
            check(lib.pc_set_small_ctas(sl.ctx, int(os.environ["PC_SMALL_CTAS"])))
    if os.environ.get("PC_TAIL"):                          # experiments: 0 = multi-kernel sequence, 1 = tail kernel on compacted atoms
        eng.set_tail(int(os.environ["PC_TAIL"]))
    n_keys = gm1.n_groups * gm2.n_groups
    # time-aggregated maps: a dense [n_keys][4] table (32 MB at most), else a hash table keyed by the 64-bit key; key spaces
    # beyond 2^32 (3 M-atom residue-pair maps: 1.1 M cells per FRAME, as many rows as a hash table has room for) send the
    # packed per-frame rows to the host instead
    rows_to_host = args.rows_to_host or n_keys >= (1 << 32)
    dense = n_keys <= (1 << 20) and not args.sparse_fold and not rows_to_host
    sparse = not dense and not rows_to_host
    large_guess = max(s1.n, n2) > 200000
    # large frames: 148 = one frame per SM for the per-frame tail kernel; 3 M-atom self contacts: the batch's accumulation
    # table lives in global memory (1.1 M (frame, key) cells per frame), 16 frames per batch keep it at ~4 GB
    batch = args.batch or ((16 if self_mode else 148) if large_guess else int(max(1, min(F, 2048, (2 << 30) // in_bytes_frame))))
    batch = min(batch, F)
    nb = (F + batch - 1) // batch
    if not args.batch:
        batch = (F + nb - 1) // nb                         # equal batches: no short last batch (1250 frames = 9 x 139, not 8 x 148 + 66)
    sparse_entries = [args.sparse_entries]
    totals = DeviceBuffer(max(1, n_keys if dense else 1) * 32, local)
    totals_t = None
    ext0 = None
    stream = eng.slots[0].stream
    if world > 1:
        ext0 = torch.cuda.ExternalStream(stream.value, device=torch.device("cuda", local))
        if dense:
            totals_t = torch.zeros(n_keys * 4, dtype=torch.float64, device="cuda")
    totals_host = _lib.PinnedArray((max(1, n_keys if dense else 1) * 4,), np.float64)
    ev = [[ctypes.c_void_p(), ctypes.c_void_p()] for _ in range(nb + 1)]
    for pair in ev:
        for e in pair:
            check(lib.pc_event_create(local, ctypes.byref(e)))
    xchg = {"cap": 0, "mine": None, "big": None}              # sparse exchange buffers (world > 1)
    ENT = _lib.fold_entry_dtype.itemsize

    def totals_ptr():
        return ctypes.c_void_p(totals_t.data_ptr()) if totals_t is not None else totals.ptr

    def exchange_sparse():
        """Ranks hold disjoint frames: every rank's sparse map -> one all-gather of fixed-size buffers on the engine's
        stream, rank 0 folds the others' entries into its table (all on the device)."""
        if xchg["cap"] == 0:
            # capacity from this rank's first table, agreed over ranks once (outside the timed region: warm-up)
            n = int(eng.fold_sparse_fetch().size)
            c = torch.tensor([n], device="cuda")
            dist.all_reduce(c, op=dist.ReduceOp.MAX)
            cap = 1024
            while cap < int(c.item()) * 3 // 2 + 1024:
                cap <<= 1
            xchg["cap"] = cap
            xchg["mine"] = torch.empty(cap * ENT, dtype=torch.uint8, device="cuda")
            xchg["big"] = torch.empty(world * cap * ENT, dtype=torch.uint8, device="cuda")
        cap = xchg["cap"]
        eng.fold_sparse_export(ctypes.c_void_p(xchg["mine"].data_ptr()), cap)
        with torch.cuda.stream(ext0):
            dist.all_gather_into_tensor(xchg["big"], xchg["mine"])
        if rank == 0:                                        # the other ranks' buffers are contiguous: one import launch
            eng.fold_sparse_import(ctypes.c_void_p(xchg["big"].data_ptr() + cap * ENT), (world - 1) * cap)

    def device_step(f_lo=0, f_hi=None, events=False, d_src=None, sort_entries=False):
        """One pass over resident frames [f_lo, f_hi) of this rank.  Returns (counts, bytes copied to the host).
        The aggregated maps arrive on the host as (key, sums) entries in table order; sort_entries orders them by key."""
        f_hi = F if f_hi is None else f_hi
        b1, b2 = d_src if d_src is not None else (d1, d2)
        tot = np.zeros(4, np.int64)
        d_rows = 0
        if dense:
            if totals_t is not None:
                with torch.cuda.stream(ext0):
                    totals_t.zero_()
            else:
                check(lib.pc_memset(local, totals.ptr, 0, n_keys * 32, stream))
        if sparse:
            eng.fold_sparse_init(sparse_entries[0])
        pending = []

        def finish(sid):
            nonlocal d_rows
            tot[:] += eng.collect_device(sid)
            if rows_to_host:
                rows, _, _ = eng.fetch_rows(sid, packed=True)      # the per-frame maps of this rank's frames -> host
                d_rows += rows.nbytes
        for b, f0 in enumerate(range(f_lo, f_hi, batch)):
            nf = min(batch, f_hi - f0)
            sid = b % len(eng.slots)
            if len(pending) == len(eng.slots):
                finish(pending.pop(0))
            eng.submit_device(sid, b1.at(f0 * n1 * 12), None if self_mode else b2.at(f0 * n2 * 12), nf,
                              accumulate=True, events=ev[b] if events else None, pack_rows=rows_to_host,
                              fold=(totals_ptr(), n_keys) if dense else ("sparse" if sparse else None))
            pending.append(sid)
        while pending:
            finish(pending.pop(0))
        if dense:
            for sl in eng.slots:
                check(lib.pc_stream_sync(sl.stream))
            if totals_t is not None:
                with torch.cuda.stream(ext0):
                    dist.all_reduce(totals_t)              # the one collective: aggregated maps over ranks
            if rank == 0:
                host = totals_host.array
                check(lib.pc_memcpy(local, host.ctypes.data, totals_ptr(), host.nbytes, 2, stream))
                d_rows += host.nbytes
            check(lib.pc_stream_sync(stream))
            device_step.last_entries = None
        elif rows_to_host:
            device_step.last_entries = None
        else:
            if dist is not None:
                exchange_sparse()
            if rank == 0 or dist is None:
                try:
                    ent = eng.fold_sparse_fetch(sort=sort_entries)   # compacted entries -> host (pinned), synchronises
                except _lib.PcError:
                    sparse_entries[0] *= 4                 # table too small: grow and redo the pass
                    xchg["cap"] = 0
                    if dist is None:
                        return device_step(f_lo, f_hi, events, d_src, sort_entries)
                    raise
                d_rows += ent.nbytes
                device_step.last_entries = ent
            else:
                check(lib.pc_stream_sync(stream))
                device_step.last_entries = None
        return tot, d_rows

    def barrier():
        if dist is not None:
            dist.barrier()
        check(lib.pc_device_sync(local))

    n_warm = max(args.warmup, 3 if not args.quick else 1)
    for k in range(n_warm):
        device_step()
        if k == 0 and sparse and not args.sparse_entries_fixed:
            # size the aggregated table from what the first pass found (x8, at least 64 k entries): clearing, compacting and
            # exchanging a 2 M-entry table every step costs ~0.2 ms, which matters once 8 ranks share the trajectory
            n_ent = int(eng.fold_sparse_fetch().size)
            if dist is not None:
                c = torch.tensor([n_ent], device="cuda")
                dist.all_reduce(c, op=dist.ReduceOp.MAX)
                n_ent = int(c.item()) * world              # rank 0 ends up with the union of all ranks' keys
            want = 1 << 16
            while want < 8 * n_ent:
                want <<= 1
            sparse_entries[0] = want
            xchg["cap"] = 0
    launches0 = lib.pc_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    barrier()
    e_all = ev[nb]
    check(lib.pc_event_record(e_all[0], stream))
    t0 = time.perf_counter()
    tot = np.zeros(4, np.int64)
    for k in range(args.steps):
        c, d_bytes = device_step()
        tot += c
    check(lib.pc_event_record(e_all[1], stream))
    ms = ctypes.c_float()
    check(lib.pc_event_elapsed_ms(e_all[0], e_all[1], ctypes.byref(ms)))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = lib.pc_launch_count() - launches0
    entries_step = None if device_step.last_entries is None else int(device_step.last_entries.size)
    if args.timeline:
        # where one step's time goes: device timestamps of every batch's scan (events on the slots' streams) relative to
        # the step's first event, host time beside them -- diagnostic only, printed to stderr
        barrier()
        check(lib.pc_event_record(e_all[0], stream))
        th0 = time.perf_counter()
        device_step(events=True)
        check(lib.pc_event_record(e_all[1], stream))
        check(lib.pc_device_sync(local))
        th1 = time.perf_counter()
        rel = ctypes.c_float()
        out = []
        for b in range(nb):
            ts = []
            for e in ev[b]:
                check(lib.pc_event_elapsed_ms(e_all[0], e, ctypes.byref(rel)))
                ts.append(rel.value)
            out.append("%d: %.3f-%.3f" % (b, ts[0], ts[1]))
        check(lib.pc_event_elapsed_ms(e_all[0], e_all[1], ctypes.byref(rel)))
        print("timeline rank %d (ms from the step's start; scan of batch b): %s | step end %.3f, host %.3f"
              % (rank, "  ".join(out), rel.value, (th1 - th0) * 1e3), file=sys.stderr)

    # ---- the dominant kernel alone: the same launches again, one batch at a time on one stream (in the step
    # above several batches are in flight, so an event pair around one launch also sees the other streams) ------
    kernel_ms = 0.0
    k_launches = 0
    large = eng_path(eng) == 2
    probe_ms = 0.0
    pe = [ctypes.c_void_p(), ctypes.c_void_p()]
    if large:
        for e in pe:
            check(lib.pc_event_create(local, ctypes.byref(e)))
        check(lib.pc_set_probe(eng.slots[0].ctx, pe[0], pe[1]))
    for k in range(2):
        for b in range(min(nb, 8)):
            f0b = b * batch
            nfb = min(batch, F - f0b)
            eng.submit_device(0, d1.at(f0b * n1 * 12), None if self_mode else d2.at(f0b * n2 * 12), nfb,
                              accumulate=False, events=ev[b])
            eng.collect_device(0)
            if nfb < batch:
                continue
            kms = ctypes.c_float()
            check(lib.pc_event_elapsed_ms(ev[b][0], ev[b][1], ctypes.byref(kms)))
            kernel_ms += kms.value
            k_launches += 1
            if large:
                check(lib.pc_event_elapsed_ms(pe[0], pe[1], ctypes.byref(kms)))
                probe_ms += kms.value
    if large:
        check(lib.pc_set_probe(eng.slots[0].ctx, None, None))
    k_launches = max(k_launches, 1)

    dev_s = max(ms.value / 1e3, 1e-9)
    if dist is not None:
        tt = torch.tensor([dev_s, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, wall = float(tt[0]), float(tt[1])
        cc = torch.tensor(tot, device="cuda")
        dist.all_reduce(cc)
        tot_all = cc.cpu().numpy()
        ff = torch.tensor([F], device="cuda")
        dist.all_reduce(ff)
        frames_step_all = int(ff.item())
    else:
        tot_all = tot
        frames_step_all = F
    value = frames_step_all * args.steps / dev_s
    contacts_step, hbonds_step, rows_step, evals_step = (tot_all / args.steps).tolist()
    step_s = dev_s / args.steps

    # ---- checked: GPU vs CPU on the frames the CPU leg analyses; E_alg of those frames ----------------------------
    names_h = np.asarray([nm.startswith("H") for nm in s.topology.names])
    heavy1 = ~names_h[s.sel1]
    heavy2 = None if self_mode else ~names_h[s.sel2]
    checked = None
    cpu = None
    ngpu_check = None
    fp32_ffma = None
    if rank == 0:
        cores = os.cpu_count() or 1
        with_cpu = world == 1 and not args.no_cpu
        from baseline import ref_runner
        use_ref = with_cpu and ref_runner.available() and REF_SAMPLE[args.workload][0] > 0
        if use_ref:
            n_serial, n_pool, _ = REF_SAMPLE[args.workload]
            chk = list(range(min(max(n_serial, n_pool), F)))
        else:
            chk = sorted(set([0, F // 2, F - 1]))
        if args.workload == "c4":
            chk = chk[:1]
        hx1 = np.empty((len(chk), n1, 3), np.float32)
        hx2 = None if self_mode else np.empty((len(chk), n2, 3), np.float32)
        for k, f in enumerate(chk):                          # exactly the bytes the scan reads
            check(lib.pc_memcpy(local, hx1[k].ctypes.data, d1.at(f * n1 * 12), n1 * 12, 2, None))
            if hx2 is not None:
                check(lib.pc_memcpy(local, hx2[k].ctypes.data, d2.at(f * n2 * 12), n2 * 12, 2, None))
        check(lib.pc_device_sync(local))
        gpu_sum = []
        for k in range(len(chk)):
            r = eng.scan(hx1[k:k + 1], None if self_mode else hx2[k:k + 1])
            gpu_sum.append((int(r.ccount[0]), int(r.hcount[0]), float(r.contacts["weight"].astype(np.float64).sum())))
        e_alg = [algorithmic_pair_evals(hx1[k], None if self_mode else hx2[k], heavy1, heavy2, CUTOFF) for k in range(min(len(chk), 3))]
        legs = None
        if use_ref:
            legs, cpu_sum = reference_loop_legs(w, hx1, hx2, cores, with_pool=True)
            by = "UNMODIFIED reference loop (baseline/_ref), nproc=1"
            n_cmp = len(cpu_sum)
        else:
            n_cmp = len(chk)
            cpu_sum = frame_summaries_oracle(w, hx1[:n_cmp], None if self_mode else hx2[:n_cmp])
            by = "oracle port (oracle/frame_oracle.py, compiled reference search)"
        ok = all(g[0] == c[0] and g[1] == c[1] and abs(g[2] - c[2]) <= 1e-5 * max(abs(c[2]), 1e-30)
                 for g, c in zip(gpu_sum[:n_cmp], cpu_sum))
        checked = {"ok": bool(ok) if n_cmp else None, "by": by, "frames": [int(lo + f) for f in chk[:n_cmp]],
                   "gpu": [list(g) for g in gpu_sum[:n_cmp]], "cpu": [list(c) for c in cpu_sum],
                   "fields": ["contacts", "hbonds", "sum_of_weights"], "tolerance": "counts exact, sum of weights 1e-5 relative"}
        if with_cpu:
            from oracle import native
            port = None
            if PORT_PER_CORE[args.workload]:
                fr = PORT_PER_CORE[args.workload] * cores
                pfps, _, _ = port_pool_run(args.workload, fr, 1, 0, cores)
                port = {"frames_per_s": pfps, "frames": fr, "nproc": cores,
                        "what": "numpy restatement of the pair loop (oracle/frame_oracle.py) around the compiled reference search"}
            if legs is not None:
                best = max([legs["nproc1"]] + ([legs["pool"]] if "pool" in legs else []), key=lambda d: d["frames_per_s"])
                cpu = {"value": best["frames_per_s"], "unit": "frames/s", "cores": best.get("nproc", 1), "kind": "reference",
                       "sample": "UNMODIFIED reference loop (baseline/_ref) on frames %d..%d of this trajectory, copied back from HBM: "
                                 "nproc=1 on %d frames, LoggingPool+chunks on %d frames; value = the faster"
                                 % (lo, lo + len(chk) - 1, legs["nproc1"]["frames"], legs.get("pool", {}).get("frames", 0)),
                       "host_cores": cores, "reference_loop": legs, "port": port}
            elif port is not None:
                cpu = {"value": port["frames_per_s"], "unit": "frames/s", "cores": cores, "kind": "port",
                       "sample": "baseline/_ref absent or no sample defined: %d frames over a %d-process pool of the oracle port" % (port["frames"], cores),
                       "host_cores": cores, "port": port}
        rate = ctypes.c_double()
        check(lib.pc_fp32_peak(local, 4096, ctypes.byref(rate)))
        fp32_ffma = rate.value
    else:
        e_alg = []

    # ---- N ranks == 1 rank on the hardware: a short trajectory sharded over the ranks vs rank 0 alone ------------
    if dist is not None:
        Fc = 4 * world
        clo, chi = worker_slices(Fc, world)[rank]
        c1b = DeviceBuffer(Fc * n1 * 12, local)
        c2b = None if self_mode else DeviceBuffer(Fc * n2 * 12, local)
        gen1.generate(c1b.ptr, 0, Fc)
        if gen2 is not None:
            gen2.generate(c2b.ptr, 0, Fc)
        check(lib.pc_device_sync(local))
        was_dense, was_sparse = dense, sparse
        dense, sparse = False, True                      # one code path for the comparison: the sparse maps
        xchg_saved = dict(xchg)
        xchg.update(cap=0, mine=None, big=None)
        c_multi, _ = device_step(clo, chi, d_src=(c1b, c2b), sort_entries=True)
        ent_multi = None if rank != 0 else device_step.last_entries.copy()
        cm = torch.tensor(c_multi[:3], device="cuda")
        dist.all_reduce(cm)
        if rank == 0:
            dist_save = dist
            dist = None                                  # rank 0 alone, no exchange
            try:
                c_single, _ = device_step(0, Fc, d_src=(c1b, c2b), sort_entries=True)
            finally:
                dist = dist_save
            ent_single = device_step.last_entries
            same_keys = ent_multi.size == ent_single.size and np.array_equal(ent_multi["key"], ent_single["key"])
            ok = same_keys and np.array_equal(cm.cpu().numpy(), c_single[:3]) and \
                np.array_equal(ent_multi["hbframes"], ent_single["hbframes"])
            rel = 0.0
            if same_keys and ent_single.size:
                for fld in ("score", "bb1", "bb2"):
                    den = np.maximum(np.abs(ent_single[fld]), 1e-30)
                    rel = max(rel, float(np.max(np.abs(ent_multi[fld] - ent_single[fld]) / den * (np.abs(ent_single[fld]) > 1e-12))))
            ngpu_check = {"ok": bool(ok and rel <= 2e-6), "frames": Fc, "ranks": world, "keys": int(ent_single.size),
                          "counts_sharded": cm.cpu().numpy().tolist(), "counts_single": c_single[:3].tolist(),
                          "max_rel_diff_sums": rel,
                          "tolerance": "keys, contact / H-bond / cell counts and frames-with-H-bond per key bit-equal; per-key sums "
                                       "2e-6 relative (per-frame cells are float32 atomic sums whose order varies run to run)"}
        dense, sparse = was_dense, was_sparse
        xchg.update(xchg_saved)
        barrier()

    # ---- roofline ---------------------------------------------------------------------------------------------------
    peaks, peak_kind = measured_peaks()
    F_all = frames_step_all
    alg_bytes_step = F_all * in_bytes_frame + 16 * contacts_step + 20 * hbonds_step + 16 * rows_step
    step_gbs = alg_bytes_step / step_s / 1e9
    k_s = kernel_ms / 1e3 / k_launches
    per_rank = dict(zip(("contacts", "hbonds", "rows", "evals"), (tot / args.steps).tolist()))
    if large and max(n1, n2) >= 200000:
        n_dom = n1 if self_mode else n2
        k_bytes = 12.0 * n_dom * batch
        k_s_dom = probe_ms / 1e3 / k_launches
        k_name = "pc_large_stream_kernel<BIN> (streaming bin of the larger selection: reads its packed coordinates once)"
    else:
        k_bytes = batch * (in_bytes_frame + (16 * per_rank["contacts"] + 20 * per_rank["hbonds"]) / max(F, 1))
        k_s_dom = k_s
        k_name = ("pc_scan_small_kernel (whole per-frame analysis in one kernel)" if not large else
                  "large-frame launch sequence of one batch (no single kernel dominates at this size)")
    traffic = traffic_step = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh).get(args.workload)
        if tj:
            traffic = tj.get("dominant_kernel_dram_bytes_per_frame", tj.get("dram_bytes_per_frame")) * batch
            if tj.get("step_dram_bytes_per_frame"):
                traffic_step = tj["step_dram_bytes_per_frame"] * F_all
    except Exception:
        pass
    e_alg_frame = float(np.mean(e_alg)) if len(e_alg) else None
    roofline = {"bound": "hbm", "achieved": step_gbs, "peak": peaks["hbm_gbs"] * world, "unit": "GB/s",
                "frac": step_gbs / (peaks["hbm_gbs"] * world), "traffic": traffic_step,
                "peak_source": peak_kind + (" x %d GPUs" % world if world > 1 else ""),
                "what": "whole step: algorithmic bytes of all frames (12 B per selected atom + 16 B per contact + 20 B per H-bond "
                        "+ 16 B per (frame, key) cell) / ms_per_step",
                "algorithmic_bytes_per_step": alg_bytes_step,
                "dominant_kernel": {"kernel": k_name, "achieved": k_bytes / k_s_dom / 1e9, "unit": "GB/s",
                                    "frac": k_bytes / k_s_dom / 1e9 / peaks["hbm_gbs"],
                                    "share_of_step": (k_s_dom * (F / batch)) / step_s, "ms_per_launch": k_s_dom * 1e3,
                                    "algorithmic_bytes_per_launch": k_bytes, "traffic": traffic,
                                    "scan_sequence_ms_per_batch": k_s * 1e3},
                "fp32": {"pair_evals_act_per_s": evals_step / step_s,
                         "pair_evals_alg_per_frame": e_alg_frame, "pair_evals_act_per_frame": evals_step / max(F_all, 1),
                         "pair_evals_alg_per_s": None if e_alg_frame is None else e_alg_frame * F_all / step_s,
                         "ffma_per_s_measured": fp32_ffma,
                         "pair_eval_peak_per_s": None if not fp32_ffma else fp32_ffma * world / 9.0,
                         "frac_alg": None if not (fp32_ffma and e_alg_frame) else e_alg_frame * F_all / step_s / (fp32_ffma * world / 9.0),
                         "what": "E_alg: reference grid, heavy atoms (src/gridsearch.C:940-985, 1111-1142), mean of the checked frames; "
                                 "E_act: candidates the kernels evaluated; peak = measured FFMA issue rate / 9 instructions per "
                                 "evaluation (3 FSUB, 3 FMUL, 2 FADD, 1 FSETP, non-fused like the reference)"}}

    # ---- e2e: the public API with host buffers (rank-local frames) ------------------------------------------------
    Fe = min(F, args.e2e_frames or 4 * batch)
    h1 = _lib.PinnedArray((Fe, n1, 3), np.float32)
    check(lib.pc_memcpy(local, h1.array.ctypes.data, d1.ptr, Fe * n1 * 12, 2, None))
    h2 = None
    if not self_mode:
        h2 = _lib.PinnedArray((Fe, n2, 3), np.float32)
        check(lib.pc_memcpy(local, h2.array.ctypes.data, d2.ptr, Fe * n2 * 12, 2, None))
    check(lib.pc_device_sync(local))
    src = ArraySource(h1.array, None if h2 is None else h2.array)
    eb = int(max(1, min(Fe, batch if large else (64 << 20) // in_bytes_frame)))

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
           "api": "ContactEngine.scan_batches (pinned host frames -> contacts, H-bonds, rows on the host), per rank"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic" if not isinstance(s, _FixtureSystem) else "reference fixture (rpn11_ubq), frames repeated",
            "config": {"workload": "%s: %s" % (args.workload, w["desc"]), "frames_per_step": F_all,
                       "frames_per_gpu_per_step": F, "trajectory_frames": F_total, "clipped": clipped,
                       "atoms_sel1": n1, "atoms_sel2": n2, "batch_frames": batch, "batches_per_gpu_per_step": nb,
                       "kernel_path": "large-frame (streaming bin of the larger selection + fused per-frame kernel, tail mode %d)" % eng.tail if large and eng.tail else
                                      "large-frame (multi-kernel sequence)" if large else "small-frame (one CTA per frame)",
                       "l2": "inputs (%.1f GB per GPU and step, read once) exceed the 126 MB L2" % (F * in_bytes_frame / 1e9),
                       "results": ("contacts and H-bonds stay in HBM; packed per-frame (frame, key) cells -> host (24 B each)" if rows_to_host else
                                   "per-frame records (contacts, H-bonds, (frame, key) cells) stay in HBM; time-aggregated maps ("
                                   + ("dense table, one all-reduce over ranks" if dense else
                                      "sparse table keyed by the 64-bit key, %s entries; ranks' tables meet in one all-gather" % entries_step)
                                   + ") copied to the host every step"),
                       "parallelism": "one trajectory cut into %d contiguous frame slices (reference chunks rule), one exchange of the aggregated maps" % world
                                      if args.scaling == "strong" else "every rank its own %d frames" % F},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "checked": checked, "ngpu_equals_1gpu": ngpu_check,
            "per_step": {"contacts": contacts_step, "hbonds": hbonds_step, "rows": rows_step,
                         "pair_evals": evals_step, "wall_ms": wall / args.steps * 1e3, "d2h_bytes": int(d_bytes)},
            "contacts_per_s": float(tot_all[0]) / dev_s, "hbonds_per_s": float(tot_all[1]) / dev_s,
            "raw_pair_evals_per_s": float(tot_all[3]) / dev_s,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def eng_path(eng):
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
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --gpus ranks share ONE trajectory (default); weak: every rank its own trajectory")
    ap.add_argument("--frames", type=int, default=0, help="frames of the trajectory (default: per workload)")
    ap.add_argument("--batch", type=int, default=0, help="frames per launch")
    ap.add_argument("--e2e-frames", type=int, default=0)
    ap.add_argument("--slots", type=int, default=3, help="frame batches in flight per GPU")
    ap.add_argument("--sparse-entries", type=int, default=1 << 21)
    ap.add_argument("--sparse-entries-fixed", action="store_true", help="keep --sparse-entries instead of sizing the table from the first pass")
    ap.add_argument("--sparse-fold", action="store_true", help="aggregated maps in the sparse table even for small key spaces")
    ap.add_argument("--rows-to-host", action="store_true", help="copy the packed per-frame rows to the host instead of aggregating on the device")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-port", action="store_true", help="--impl reference: skip the port's figure")
    ap.add_argument("--quick", action="store_true", help="profiling runs: 1 warm-up step")
    ap.add_argument("--timeline", action="store_true", help="diagnostic: device timestamps of one step's batches on stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
