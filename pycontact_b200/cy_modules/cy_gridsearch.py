"""The reference's Cython entry points (PyContact/cy_modules/cy_gridsearch.pyx) on the CUDA library:

``cy_find_contacts`` (:31-35 -> find_contacts, cy_modules/src/gridsearch.C:408-427) -> ``pc_find_contacts``;
``cy_sasa`` (:14-22 -> sasa_grid, gridsearch.C:430-529) -> ``pc_sasa``;
``cy_find_within`` (:24-29 -> find_within, gridsearch.C:595-713) -> ``pc_find_within``;
plus the frame-batched forms ``sasa_frames`` / ``find_within_frames`` that the host code uses.

``cy_find_contacts`` keeps signature and return shape: a list of ``natoms1 + natoms2`` lists of ints, the first
``natoms1`` holding, for every atom of set 1, the set-2 indices within ``r`` in the reference's
visiting order, the last ``natoms2`` always empty (gridsearch.C:410, 421-422).
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib


def find_contacts_csr(xyz1, natoms1, xyz2, natoms2, r, device=0):
    """CSR form (row_ptr int64[natoms1+1], cols int32) -- what the object-free callers use."""
    a = np.asarray(xyz1)
    b = np.asarray(xyz2)
    if a.dtype != np.float32 or b.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float'")     # Cython's error for float[:, :]
    if not a.flags["C_CONTIGUOUS"] or not b.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")
    natoms1, natoms2 = int(natoms1), int(natoms2)
    if a.size < 3 * natoms1 or b.size < 3 * natoms2:
        raise ValueError("coordinate arrays are shorter than 3 * natoms")
    lib = _lib.load()
    row_ptr = np.zeros(natoms1 + 1, np.int64)
    cols_p = ctypes.c_void_p()
    _lib.check(lib.pc_find_contacts(int(device), _lib.ptr(a), natoms1, _lib.ptr(b), natoms2, float(r),
                                    _lib.ptr(row_ptr), ctypes.byref(cols_p)))
    total = int(row_ptr[-1])
    try:
        if total:
            cols = np.ctypeslib.as_array(ctypes.cast(cols_p, ctypes.POINTER(ctypes.c_int32)), shape=(total,)).copy()
        else:
            cols = np.zeros(0, np.int32)
    finally:
        if cols_p.value:
            lib.pc_free_host(cols_p)
    return row_ptr, cols


def cy_find_contacts(xyz1, natoms1, xyz2, natoms2, r):
    row_ptr, cols = find_contacts_csr(xyz1, natoms1, xyz2, natoms2, r)
    flat = cols.tolist()
    bounds = row_ptr.tolist()
    out = [flat[bounds[i]:bounds[i + 1]] for i in range(int(natoms1))]
    out.extend([] for _ in range(int(natoms2)))
    return out


def _f32_rows(a, what):
    a = np.asarray(a)
    if a.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float' but got %r (%s)" % (a.dtype.name, what))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")
    return a


def _i32(a, what):
    a = np.asarray(a)
    if a.dtype != np.int32:
        raise ValueError("Buffer dtype mismatch, expected 'int' but got %r (%s)" % (a.dtype.name, what))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")
    return a


def sasa_frames(coords, radius, pairdist, npts, probe_radius, pointstyle, restricted, restricted_list,
                device=0, return_exposed=False):
    """SASA of ``F`` frames of one selection in one device call: ``coords`` float32 (F, natoms, 3).
    Returns float32[F] (the reference's per-frame float total, bit for bit) and, on request, the exposed
    sphere points per atom (F, natoms)."""
    c = np.ascontiguousarray(coords, np.float32)
    if c.ndim != 3 or c.shape[2] != 3:
        raise ValueError("coords must be (frames, natoms, 3)")
    F, n = int(c.shape[0]), int(c.shape[1])
    rad = np.ascontiguousarray(radius, np.float32)
    if rad.size < n:
        raise ValueError("radius array is shorter than natoms")
    rl = np.ascontiguousarray(restricted_list, np.int32)
    if restricted and rl.size < n:
        raise ValueError("restrictedList is shorter than natoms")
    total = np.zeros(F, np.float32)
    exposed = np.zeros((F, n), np.int32) if return_exposed else None
    lib = _lib.load()
    _lib.check(lib.pc_sasa(int(device), _lib.ptr(c), F, n, _lib.ptr(rad), float(pairdist), int(npts),
                           float(probe_radius), int(pointstyle), int(bool(restricted)), _lib.ptr(rl),
                           _lib.ptr(total), _lib.ptr(exposed)))
    return (total, exposed) if return_exposed else total


def cy_sasa(npcoords, natoms, pairdist, allow_double_counting, maxsize, nprad, surfacePoints, probeRadius,
            pointstyle, restricted, restrictedList):
    """Same arguments as the reference; ``allow_double_counting`` / ``maxsize`` are ignored there too (the
    wrapper passes 0 and -1).  Returns a Python float holding the C float."""
    c = _f32_rows(npcoords, "npcoords")
    rad = _f32_rows(nprad, "nprad")
    rl = _i32(restrictedList, "restrictedList")
    natoms = int(natoms)
    if c.size < 3 * natoms or rad.size < natoms:
        raise ValueError("coordinate / radius arrays are shorter than natoms")
    tot = sasa_frames(c.reshape(-1)[:3 * natoms].reshape(1, natoms, 3), rad, pairdist, surfacePoints, probeRadius,
                      pointstyle, restricted, rl)
    return float(tot[0])


def find_within_frames(coords, flgs, others, r, device=0):
    """``find_within`` for (F, num, 3) float32 frames with per-atom 0/1 ``flgs`` / ``others``: int32 (F, num)."""
    c = np.ascontiguousarray(coords, np.float32)
    if c.ndim != 3 or c.shape[2] != 3:
        raise ValueError("coords must be (frames, num, 3)")
    F, n = int(c.shape[0]), int(c.shape[1])
    f = np.ascontiguousarray(flgs, np.int32)
    o = np.ascontiguousarray(others, np.int32)
    if f.size < n or o.size < n:
        raise ValueError("flag arrays are shorter than num")
    res = np.zeros((F, n), np.int32)
    lib = _lib.load()
    _lib.check(lib.pc_find_within(int(device), _lib.ptr(c), F, n, _lib.ptr(f), _lib.ptr(o), float(r), _lib.ptr(res)))
    return res


def cy_find_within(xyz, flgs, others, num, r):
    """Returns int32[num]; like the reference, ``flgs`` is overwritten with the result
    (gridsearch.C:675, 700-707 clear it and OR the result back in)."""
    c = _f32_rows(xyz, "xyz")
    f = _i32(flgs, "flgs")
    o = _i32(others, "others")
    num = int(num)
    res = find_within_frames(c.reshape(-1)[:3 * num].reshape(1, num, 3), f[:num], o[:num], r)[0]
    if f.flags.writeable:
        f[:num] = res
    return res
