"""``cy_find_contacts`` -- compatibility shim for the reference's Cython entry point
(PyContact/cy_modules/cy_gridsearch.pyx:31-35 -> find_contacts,
cy_modules/src/gridsearch.C:408-427) backed by ``pc_find_contacts`` of the CUDA library.

Same signature, same return shape: a list of ``natoms1 + natoms2`` lists of ints, the first
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
