"""Synthetic MD-like systems for the benchmark configurations.

There is no network and the reference ships one tiny fixture, so the named
benchmark sizes are generated: protein-like residues of 16 atoms (8 heavy, 8 H),
3-site water, lipid-like residues of 130 atoms, at ~0.1 atoms/A^3, no periodic
boundaries (the reference ignores the unit cell, App. A.6 Q5).

``make_system(name, seed)`` builds topology + frame-0 coordinates on the host
(numpy).  Frames are frame 0 + a rigid per-residue oscillation + per-atom Gaussian
jitter; ``frames_host`` is the numpy generator used by CPU tests and golden vectors,
the device generator (``pc_synth_frames`` in the CUDA library) is used by bench.py --
inputs are *defined* as the float32 values that end up in memory, and parity is always
checked on exactly those bytes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict

import numpy as np

from .io.psf import Topology

# residue templates: (name, type, parent index or -1, bond length)
_SIDE_END = [("OG", "OH1", "HG", "H"), ("NZ", "NH3", "HZ1", "HC"), ("SD", "S", "HD9", "HS"), ("CE", "CT3", "HE1", "HA3")]
_RESNAMES = ["SER", "LYS", "MET", "LEU", "ASP", "GLU", "THR", "VAL", "ALA", "ILE", "ASN", "GLN"]
_PROT_HEAVY = [("N", "NH1"), ("CA", "CT1"), ("C", "C"), ("O", "O"), ("CB", "CT2"), ("CG", "CT2"), ("CD", "CT2")]
_PROT_H = [("HN", "H", 0, 1.01), ("HA", "HB1", 1, 1.09), ("HB1", "HA2", 4, 1.09), ("HB2", "HA2", 4, 1.09),
           ("HG1", "HA2", 5, 1.09), ("HG2", "HA2", 5, 1.09), ("HD1", "HA2", 6, 1.09)]
RES_ATOMS = 16
LIPID_ATOMS = 130
LIPID_HEAVY = 50


@dataclass
class SynthSystem:
    name: str
    seed: int
    topology: Topology
    coords0: np.ndarray          # float32 (natoms, 3) frame 0
    residue_of: np.ndarray       # int32 (natoms,) residue/molecule ordinal (for rigid motion)
    amp: np.ndarray              # float32 (nres, 3) oscillation amplitude per residue
    period: np.ndarray           # float32 (nres,)
    phase: np.ndarray            # float32 (nres,)
    sel1: np.ndarray
    sel2: np.ndarray
    sel1text: str
    sel2text: str
    backbone: np.ndarray
    maps: tuple = ([0, 0, 1, 1, 0], [0, 0, 1, 1, 0])
    n_frames: int = 10000
    sigma: float = 0.25
    meta: Dict = field(default_factory=dict)


def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v


def _snake_sites(nx, ny, nz, spacing, origin):
    """lattice sites in boustrophedon order so consecutive residues are neighbours"""
    pts = []
    for iz in range(nz):
        ys = range(ny) if iz % 2 == 0 else range(ny - 1, -1, -1)
        for n_y, iy in enumerate(ys):
            xs = range(nx) if (n_y + iz * ny) % 2 == 0 else range(nx - 1, -1, -1)
            for ix in xs:
                pts.append((ix, iy, iz))
    return np.asarray(pts, np.float64) * spacing + np.asarray(origin, np.float64)


class _Builder:
    def __init__(self, rng):
        self.rng = rng
        self.seg, self.resid, self.resname, self.name, self.type = [], [], [], [], []
        self.xyz, self.bonds, self.res_of = [], [], []
        self.n = 0
        self.nres = 0

    def add_protein(self, centres, segprefix, seg_len=400):
        rng = self.rng
        nres = len(centres)
        kinds = rng.integers(0, len(_SIDE_END), nres)
        rnames = rng.integers(0, len(_RESNAMES), nres)
        heavy_off = _unit(rng, nres * 8).reshape(nres, 8, 3) * rng.uniform(0.6, 2.4, (nres, 8, 1))
        hdir = _unit(rng, nres * 8).reshape(nres, 8, 3)
        first = self.n
        for r in range(nres):
            base = self.n
            seg = "%s%d" % (segprefix, r // seg_len)
            rid = r % seg_len + 1
            end = _SIDE_END[kinds[r]]
            heavy = _PROT_HEAVY + [(end[0], end[1])]
            hs = _PROT_H + [(end[2], end[3], 7, 0.96 if end[0] == "OG" else 1.01 if end[0] == "NZ" else 1.09)]
            # interleave like a real PSF: each heavy atom followed by its hydrogens
            order = []
            for hi in range(8):
                order.append(("heavy", hi))
                for k, h in enumerate(hs):
                    if h[2] == hi:
                        order.append(("h", k))
            pos_in_res = {}
            hpos = {}
            for slot, (kind, k) in enumerate(order):
                pos_in_res[(kind, k)] = base + slot
            for kind, k in order:
                if kind == "heavy":
                    self.name.append(heavy[k][0]); self.type.append(heavy[k][1])
                    self.xyz.append(centres[r] + heavy_off[r, k])
                else:
                    h = hs[k]
                    self.name.append(h[0]); self.type.append(h[1])
                    self.xyz.append(centres[r] + heavy_off[r, h[2]] + hdir[r, k] * h[3])
                    self.bonds.append((pos_in_res[("heavy", h[2])], pos_in_res[("h", k)]))
                self.seg.append(seg); self.resid.append(rid); self.resname.append(_RESNAMES[rnames[r]])
                self.res_of.append(self.nres)
            # heavy-heavy bonds (N-CA, CA-C, C-O, CA-CB, CB-CG, CG-CD, CD-X)
            for a, b in ((0, 1), (1, 2), (2, 3), (1, 4), (4, 5), (5, 6), (6, 7)):
                self.bonds.append((pos_in_res[("heavy", a)], pos_in_res[("heavy", b)]))
            self.n += RES_ATOMS
            self.nres += 1
        return np.arange(first, self.n)

    def add_water(self, centres, seg="WT"):
        rng = self.rng
        nw = len(centres)
        first = self.n
        d1 = _unit(rng, nw)
        t = _unit(rng, nw)
        t -= (t * d1).sum(1, keepdims=True) * d1
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        ang = np.deg2rad(104.52)
        d2 = np.cos(ang) * d1 + np.sin(ang) * t
        o = np.asarray(centres)
        h1 = o + 0.9572 * d1
        h2 = o + 0.9572 * d2
        xyz = np.stack([o, h1, h2], axis=1).reshape(-1, 3)
        self.xyz.extend(xyz)
        for w in range(nw):
            b = self.n
            self.name += ["OH2", "H1", "H2"]; self.type += ["OT", "HT", "HT"]
            self.seg += ["%s%d" % (seg, w // 9999)] * 3
            self.resid += [w % 9999 + 1] * 3
            self.resname += ["TIP3"] * 3
            self.res_of += [self.nres] * 3
            self.bonds += [(b, b + 1), (b, b + 2), (b + 1, b + 2)]
            self.n += 3
            self.nres += 1
        return np.arange(first, self.n)

    def add_lipids(self, centres, seg="MEM"):
        rng = self.rng
        first = self.n
        for li, c in enumerate(centres):
            base = self.n
            # 50 heavy atoms in a cylinder r<=4.5, |z|<=10 ; 80 H bonded round-robin to carbons
            rr = 4.5 * np.sqrt(rng.uniform(0, 1, LIPID_HEAVY))
            th = rng.uniform(0, 2 * np.pi, LIPID_HEAVY)
            zz = np.linspace(-10, 10, LIPID_HEAVY)
            hv = np.stack([rr * np.cos(th), rr * np.sin(th), zz], 1) + c
            hnames = ["P"] + ["O1%d" % k for k in range(1, 9)] + ["N"] + ["C%d" % k for k in range(1, 41)]
            htypes = ["PL"] + ["O2L"] * 8 + ["NTL"] + ["CTL2"] * 40
            carbons = list(range(10, 50))
            # PSF-like interleaving: heavy atom followed by its hydrogens
            per = [[] for _ in range(LIPID_HEAVY)]
            for k in range(LIPID_ATOMS - LIPID_HEAVY):
                per[carbons[k % 40]].append(k)
            hd = _unit(rng, LIPID_ATOMS - LIPID_HEAVY)
            idx_heavy = []
            for a in range(LIPID_HEAVY):
                idx_heavy.append(self.n)
                self.name.append(hnames[a]); self.type.append(htypes[a]); self.xyz.append(hv[a])
                self.n += 1
                for k in per[a]:
                    self.name.append("H%dA" % k if k < 100 else "H%d" % k); self.type.append("HAL2")
                    self.xyz.append(hv[a] + 1.09 * hd[k])
                    self.bonds.append((idx_heavy[a], self.n))
                    self.n += 1
            for a in range(LIPID_HEAVY - 1):
                self.bonds.append((idx_heavy[a], idx_heavy[a + 1]))
            self.seg += [seg] * LIPID_ATOMS
            self.resid += [li + 1] * LIPID_ATOMS
            self.resname += ["POPC"] * LIPID_ATOMS
            self.res_of += [self.nres] * LIPID_ATOMS
            self.nres += 1
            assert self.n - base == LIPID_ATOMS
        return np.arange(first, self.n)

    def topology(self):
        n = self.n
        return Topology(self.seg, np.asarray(self.resid, np.int64), self.resname, self.name, self.type,
                        np.zeros(n), np.ones(n), np.asarray(self.bonds, np.int64).reshape(-1, 2))


def _water_sites(rng, lo, hi, spacing, exclude, n):
    """n jittered lattice sites in the box [lo,hi] not inside any exclusion box"""
    ax = [np.arange(lo[k] + spacing / 2, hi[k], spacing) for k in range(3)]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    ok = np.ones(len(g), bool)
    for elo, ehi in exclude:
        ok &= ~np.all((g >= np.asarray(elo)) & (g <= np.asarray(ehi)), axis=1)
    g = g[ok]
    if len(g) < n:
        raise ValueError("box too small for %d waters (%d sites)" % (n, len(g)))
    # keep the sites closest to the box centre so the solvent shell is compact
    ctr = (np.asarray(lo) + np.asarray(hi)) / 2
    order = np.argsort(np.abs(g - ctr).max(axis=1), kind="stable")
    g = g[order[:n]]
    return g + rng.uniform(-0.4, 0.4, g.shape)


SPACING = 5.43            # (16 atoms / 0.1 A^-3)^(1/3)


def _finish(name, seed, b, sel1, sel2, s1t, s2t, rng, n_frames, maps=None, meta=None):
    topo = b.topology()
    coords0 = np.asarray(b.xyz, np.float32).reshape(-1, 3)
    nres = b.nres
    amp = (rng.uniform(0, 1.5, (nres, 1)) * _unit(rng, nres)).astype(np.float32)
    period = rng.uniform(50, 400, nres).astype(np.float32)
    phase = rng.uniform(0, 2 * np.pi, nres).astype(np.float32)
    from .io.selection import select_atoms
    backbone = select_atoms(topo, "backbone")
    s = SynthSystem(name, seed, topo, coords0, np.asarray(b.res_of, np.int32), amp, period, phase,
                    np.asarray(sel1, np.int64), np.asarray(sel2, np.int64), s1t, s2t, backbone,
                    n_frames=n_frames, meta=meta or {})
    if maps is not None:
        s.maps = maps
    return s


def _two_chain(name, seed, nres_chain, n_water, n_frames, interleave=False):
    rng = np.random.default_rng(seed)
    b = _Builder(rng)
    nx = max(1, int(round(nres_chain ** (1 / 3) * 0.6)))
    ny = int(np.ceil(np.sqrt(nres_chain / nx)))
    nz = int(np.ceil(nres_chain / (nx * ny)))
    if interleave:
        # both chains on ONE lattice, residues alternating: the selections' bounding boxes coincide and every
        # atom has partners (the case the bounding-box prefilter of the scan kernels cannot thin out)
        both = _snake_sites(2 * nx, ny, nz, SPACING, (-nx * SPACING + SPACING / 2, 0, 0))[:2 * nres_chain]
        sa, sb = both[0::2], both[1::2]
    else:
        sa = _snake_sites(nx, ny, nz, SPACING, (-nx * SPACING + SPACING / 2, 0, 0))[:nres_chain]
        sb = _snake_sites(nx, ny, nz, SPACING, (SPACING / 2, 0, 0))[:nres_chain]
    a = b.add_protein(sa, "PA")
    c = b.add_protein(sb, "PB")
    lo = np.minimum(sa.min(0), sb.min(0)) - 3.0
    hi = np.maximum(sa.max(0), sb.max(0)) + 3.0
    if n_water:
        pad = 3.1 * (np.ceil((n_water / 0.0334) ** (1 / 3) / 3.1) + 2)
        ctr = (lo + hi) / 2
        half = np.maximum((hi - lo) / 2 + 6.0, pad / 2)
        w = _water_sites(rng, ctr - half * 1.6, ctr + half * 1.6, 3.1, [(lo, hi)], n_water)
        b.add_water(w)
    return _finish(name, seed, b, a, c, "segid PA*", "segid PB*", rng, n_frames,
                   meta={"nres_chain": nres_chain, "n_water": n_water})


def make_system(name: str, seed: int = None) -> SynthSystem:
    """name in {"tiny", "small", "small_i", "c2", "c2i", "c3", "c3_small", "c4", "c4_small"}  (SURVEY.md §8 configs)."""
    if name == "tiny":          # CPU tests / golden vectors: 2 x 24 residues + 120 waters = 1128 atoms
        return _two_chain(name, 7 if seed is None else seed, 24, 120, 6)
    if name == "small":         # GPU parity at oracle-friendly size: 2 x 120 residues + 1000 waters
        return _two_chain(name, 11 if seed is None else seed, 120, 1000, 64)
    if name == "c2":            # 100 032 atoms: 2 x 375 residues (6000 atoms each) + 29 344 waters
        return _two_chain(name, 1002 if seed is None else seed, 375, 29344, 10000)
    if name == "c2i":           # c2 with the two chains interdigitated (every selected atom inside the search grid)
        return _two_chain(name, 1002 if seed is None else seed, 375, 29344, 10000, interleave=True)
    if name == "small_i":       # the same at test size
        return _two_chain(name, 11 if seed is None else seed, 120, 1000, 64, interleave=True)
    if name in ("c3", "c3_small"):
        seed = 1003 if seed is None else seed
        rng = np.random.default_rng(seed)
        b = _Builder(rng)
        if name == "c3":        # 999 998 atoms: 3125 residues + 3000 lipids + 186 666 waters
            nres, nlip, nwat, side = 3125, 3000, 186666, (15, 15, 14)
        else:                   # 1/20 scale, same composition (GPU parity against the oracle)
            nres, nlip, nwat, side = 156, 150, 9333, (6, 6, 5)
        sp = _snake_sites(side[0], side[1], side[2], SPACING, (0, 0, 0))[:nres]
        sp -= sp.mean(0)
        p = b.add_protein(sp, "PR")
        plo, phi = sp.min(0) - 3.5, sp.max(0) + 3.5
        # lipids: upright cylinders (9 A pitch) in a slab |z|<=10 around the protein's xy footprint
        nside = int(np.ceil(np.sqrt(nlip + ((phi[0] - plo[0]) / 9.0 + 1) * ((phi[1] - plo[1]) / 9.0 + 1)))) + 2
        gx = (np.arange(nside) - nside / 2 + 0.5) * 9.0
        lg = np.stack(np.meshgrid(gx, gx, indexing="ij"), -1).reshape(-1, 2)
        outside = ~((lg[:, 0] >= plo[0] - 4.5) & (lg[:, 0] <= phi[0] + 4.5) &
                    (lg[:, 1] >= plo[1] - 4.5) & (lg[:, 1] <= phi[1] + 4.5))
        lg = lg[outside]
        lg = lg[np.argsort(np.abs(lg).max(1), kind="stable")][:nlip]
        assert len(lg) == nlip
        lc = np.concatenate([lg, np.zeros((nlip, 1))], 1)
        l = b.add_lipids(lc)
        mlo = np.array([lg[:, 0].min() - 4.5, lg[:, 1].min() - 4.5, -10.5])
        mhi = np.array([lg[:, 0].max() + 4.5, lg[:, 1].max() + 4.5, 10.5])
        half = max(np.abs(mlo).max(), np.abs(mhi).max(), (nwat / 0.0334) ** (1 / 3) / 2) * 1.35
        w = b.add_water(_water_sites(rng, (-half,) * 3, (half,) * 3, 3.1, [(plo, phi), (mlo, mhi)], nwat))
        sel2 = np.concatenate([l, w])
        return _finish(name, seed, b, p, sel2, "segid PR*", "not segid PR*", rng, 10000,
                       meta={"nres": nres, "nlip": nlip, "nwat": nwat})
    if name in ("c4", "c4_small"):
        seed = 1004 if seed is None else seed
        rng = np.random.default_rng(seed)
        b = _Builder(rng)
        nres = 187500 if name == "c4" else 4000
        thick = 30.0
        rmid = np.sqrt(nres * 160.0 / (4 * np.pi * thick))
        rout = rmid + thick / 2 + SPACING
        ax = np.arange(-rout, rout + SPACING, SPACING)
        # build shell sites slab by slab to bound memory
        pts = []
        for z in ax:
            g = np.stack(np.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2)
            r = np.sqrt((g ** 2).sum(1) + z * z)
            m = np.abs(r - rmid) <= thick / 2 + 0.35 * SPACING
            if m.any():
                gg = g[m]
                if int(round((z + rout) / SPACING)) % 2:
                    gg = gg[::-1]
                pts.append(np.concatenate([gg, np.full((len(gg), 1), z)], 1))
        pts = np.concatenate(pts)
        if len(pts) < nres:
            raise ValueError("shell lattice too small: %d < %d" % (len(pts), nres))
        keep = np.sort(np.argsort(np.abs(np.linalg.norm(pts, axis=1) - rmid), kind="stable")[:nres])
        p = b.add_protein(pts[keep], "CP")
        return _finish(name, seed, b, p, p, "segid CP*", "self", rng, 10000,
                       maps=([0, 0, 1, 1, 1], [0, 0, 1, 1, 1]), meta={"nres": nres, "rmid": float(rmid)})
    raise ValueError("unknown synthetic system %r" % name)


def cloud(n1, n2, density=0.1, heavy_fraction=0.5, seed=1005):
    """C5: two uniform random clouds sharing one cubic box (sweep config); returns
    (xyz1, xyz2, heavy1, heavy2) with float32 coords and boolean heavy masks."""
    rng = np.random.default_rng(seed)
    edge = ((n1 + n2) / density) ** (1 / 3)
    a = rng.uniform(0, edge, (n1, 3)).astype(np.float32)
    b = rng.uniform(0, edge, (n2, 3)).astype(np.float32)
    return a, b, rng.uniform(size=n1) < heavy_fraction, rng.uniform(size=n2) < heavy_fraction


def frames_host(system: SynthSystem, frames, atoms=None) -> np.ndarray:
    """float32 (len(frames), n, 3) coordinates (numpy generator; see module docstring)."""
    atoms = np.arange(system.topology.n_atoms) if atoms is None else np.asarray(atoms)
    res = system.residue_of[atoms]
    base = system.coords0[atoms].astype(np.float64)
    out = np.empty((len(frames), len(atoms), 3), np.float32)
    for k, f in enumerate(frames):
        rng = np.random.default_rng([system.seed, int(f)])
        jit = rng.normal(0.0, system.sigma, (system.topology.n_atoms, 3))[atoms]
        osc = system.amp[res] * np.sin(2 * np.pi * f / system.period[res] + system.phase[res])[:, None]
        out[k] = (base + osc + jit).astype(np.float32)
    return out
