"""oracle/sasa_oracle.py -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's solvent-accessible-surface and ``within`` routines:

  sasa_grid     PyContact/cy_modules/src/gridsearch.C:430-529 (Shrake-Rupley with shared sphere points;
                neighbour list from vmd_gridsearch1 :208-407: 0.001 <= distance2 <= pairdist^2)
  find_within   :595-713 with find_within_routine :715-826

float32 arithmetic in the reference's order of operations (numpy float32 scalars, no fused operations).
The sphere points are drawn exactly like the reference draws them: libc ``srandom(38572111)`` /
``random()`` and libm's ``cosf`` / ``sinf`` / ``sqrtf`` / ``acosf`` through ctypes.
Pinned against the UNMODIFIED reference compiled into oracle/_ref (``ref_sasa`` / ``ref_find_within`` below,
tests/test_sasa_cpu.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import ctypes
import ctypes.util
import math
import os

import numpy as np

f32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "libref_gridsearch.so")


def _libs():
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libc.random.restype = ctypes.c_long
    libc.srandom.argtypes = [ctypes.c_uint]
    for fn in ("cosf", "sinf", "sqrtf", "acosf"):
        getattr(libm, fn).restype = ctypes.c_float
        getattr(libm, fn).argtypes = [ctypes.c_float]
    return libc, libm


def sphere_points(npts, pointstyle):
    """float32 [npts, 3], gridsearch.C:453-484."""
    libc, libm = _libs()
    pts = np.zeros((npts, 3), f32)
    if pointstyle:
        inv = f32(1.0) / f32(2147483647)
        libc.srandom(38572111)
        for i in range(npts):
            u1 = f32(libc.random())
            u2 = f32(libc.random())
            z = f32(f32(f32(2.0) * u1) * inv) - f32(1.0)
            phi = f32(2.0 * math.pi * float(u2) * float(inv))
            R = f32(libm.sqrtf(f32(1.0) - f32(z * z)))
            pts[i] = (f32(R * f32(libm.cosf(phi))), f32(R * f32(libm.sinf(phi))), z)
    else:
        for k in range(npts):
            with np.errstate(invalid="ignore", divide="ignore"):
                h = f32(np.float64(2.0 * (k - 1.0)) / np.float64(npts - 1.0) - 1.0)
            theta = f32(libm.acosf(h))
            phi = f32(0)
            if not (k == 1 or k == npts):
                # sqrt(npts * (1 - (h_k*h_k))) has a float argument, so <math.h> under `using namespace std`
                # picks the float overload (:476); the division and the fmod are double
                with np.errstate(invalid="ignore", divide="ignore"):
                    root = float(libm.sqrtf(f32(f32(npts) * f32(f32(1.0) - f32(h * h)))))
                    phi = f32(math.fmod(float(phi) + 3.6 / root, 2 * math.pi)) if root == root else f32(np.nan)
            st = f32(libm.sinf(theta))
            pts[k] = (f32(f32(libm.cosf(phi)) * st), f32(f32(libm.sinf(phi)) * st), f32(libm.cosf(theta)))
    return pts


def _d2(a, b):
    d = f32(a[0] - b[0]); r2 = f32(d * d)
    d = f32(a[1] - b[1]); r2 = f32(r2 + f32(d * d))
    d = f32(a[2] - b[2])
    return f32(r2 + f32(d * d))


def sasa(pos, radius, pairdist, npts, srad, pointstyle, restricted, restricted_list):
    """-> (total area as the reference's float32 running sum, per-atom areas float32, per-atom exposed points)."""
    pos = np.asarray(pos, f32).reshape(-1, 3)
    n = len(pos)
    radius = np.asarray(radius, f32)
    pts = sphere_points(npts, pointstyle)
    sq = f32(f32(pairdist) * f32(pairdist))
    prefac = f32(4 * math.pi / npts)
    radp = np.array([f32(float(r) + float(srad)) for r in radius], f32)
    total = f32(0)
    areas = np.zeros(n, f32)
    exposed = np.zeros(n, np.int32)
    # all-pairs distance2 in float32, reference association order
    d = pos[:, None, :] - pos[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    nb = (d2.astype(np.float64) >= 0.001) & (d2 <= sq)
    np.fill_diagonal(nb, False)
    for i in range(n):
        if restricted and not restricted_list[i]:
            continue
        rad = radp[i]
        nbrs = np.nonzero(nb[i])[0]
        surf = pos[i][None, :] + rad * pts                    # float32: loc + rad * spherept
        if len(nbrs):
            dd = surf[:, None, :] - pos[nbrs][None, :, :]
            dist = (dd[..., 0] * dd[..., 0] + dd[..., 1] * dd[..., 1]) + dd[..., 2] * dd[..., 2]
            rsq = radp[nbrs] * radp[nbrs]
            buried = (dist <= rsq[None, :]).any(axis=1)
        else:
            buried = np.zeros(npts, bool)
        cnt = int(npts - buried.sum())
        exposed[i] = cnt
        area = f32(f32(f32(prefac * rad) * rad) * f32(cnt))
        areas[i] = area
        total = f32(total + area)
    return total, areas, exposed


def find_within(xyz, flgs, others, r):
    """int32 [num]: flagged atoms within r (strictly, float32) of an ``others`` atom."""
    xyz = np.asarray(xyz, f32).reshape(-1, 3)
    flgs = np.asarray(flgs).astype(bool)
    others = np.asarray(others).astype(bool)
    res = np.zeros(len(xyz), np.int32)
    if not flgs.any() or not others.any():
        return res
    r2 = f32(f32(r) * f32(r))
    o = xyz[others]
    for i in np.nonzero(flgs)[0]:
        d = xyz[i][None, :] - o
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        res[i] = int((d2 < r2).any())
    return res


# ---- the UNMODIFIED reference, compiled by oracle/Makefile ------------------------------------------------
def have_ref():
    return os.path.exists(_REF_SO)


def _ref():
    lib = ctypes.CDLL(_REF_SO)
    lib.ref_sasa_grid.restype = ctypes.c_double
    lib.ref_sasa_grid.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int,
                                  ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.ref_find_within.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_void_p]
    return lib


def ref_sasa(pos, radius, pairdist, npts, srad, pointstyle, restricted, restricted_list):
    pos = np.ascontiguousarray(pos, f32).reshape(-1)
    radius = np.ascontiguousarray(radius, f32)
    rl = np.ascontiguousarray(restricted_list, np.int32)
    return f32(_ref().ref_sasa_grid(pos.ctypes.data, len(radius), float(pairdist), radius.ctypes.data, int(npts),
                                    float(srad), int(pointstyle), int(restricted), rl.ctypes.data))


def ref_find_within(xyz, flgs, others, r):
    xyz = np.ascontiguousarray(xyz, f32).reshape(-1)
    f = np.ascontiguousarray(flgs, np.int32).copy()
    o = np.ascontiguousarray(others, np.int32).copy()
    res = np.zeros(len(f), np.int32)
    _ref().ref_find_within(xyz.ctypes.data, f.ctypes.data, o.ctypes.data, len(f), float(r), res.ctypes.data)
    return res
