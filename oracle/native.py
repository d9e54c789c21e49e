"""oracle/native.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers of the native neighbour search:

* ``oracle_find_contacts``  -- our C restatement (oracle/gridsearch_oracle.c) of
  ``find_contacts`` / ``vmd_gridsearch3`` (PyContact/cy_modules/src/gridsearch.C:408-427, 902-1162)
* ``ref_find_contacts``     -- the UNMODIFIED reference gridsearch.C compiled into
  ``oracle/_ref/libref_gridsearch.so`` by oracle/Makefile (kind "reference").

Both return ``(row_ptr int64[n1+1], cols int32[P])``: for every atom of set 1 the
set-2 neighbours in the reference's visiting order.  ``as_list_of_lists`` gives the
exact return shape of ``cy_find_contacts`` (cy_gridsearch.pyx:31-35): n1+n2 rows,
the last n2 empty.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import resource
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_gridsearch.so")

_oracle = None
_ref = None


def build(force: bool = False) -> None:
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    if force or not os.path.exists(_ORACLE_SO) or \
            os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "gridsearch_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "liboracle.so")])
    shim = os.path.join(_HERE, "ref_shim.cpp")
    if force or not os.path.exists(_REF_SO) or os.path.getmtime(_REF_SO) < os.path.getmtime(shim):
        # the Makefile keeps a prebuilt _ref when /root/reference is absent (the GPU box)
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _load_oracle():
    global _oracle
    if _oracle is None:
        build()
        lib = ctypes.CDLL(_ORACLE_SO)
        lib.oracle_find_contacts.restype = ctypes.c_int
        lib.oracle_find_contacts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_double, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.oracle_free.argtypes = [ctypes.c_void_p]
        lib.oracle_grid_geometry.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_box_coords.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_float,
                                          ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_stencil_rank.restype = ctypes.c_int
        lib.oracle_stencil_rank.argtypes = [ctypes.c_int] * 3
        lib.oracle_brute_count.restype = ctypes.c_int64
        lib.oracle_brute_count.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_void_p]
        _oracle = lib
    return _oracle


def have_reference() -> bool:
    return os.path.exists(_REF_SO)


def _load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            build()
        if not os.path.exists(_REF_SO):
            raise RuntimeError("oracle/_ref/libref_gridsearch.so is not built (no /root/reference here)")
        lib = ctypes.CDLL(_REF_SO)
        lib.ref_find_contacts_csr.restype = ctypes.c_int
        lib.ref_find_contacts_csr.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_double, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.ref_free.argtypes = [ctypes.c_void_p]
        _ref = lib
    return _ref


def _prep(xyz):
    a = np.ascontiguousarray(np.asarray(xyz, dtype=np.float32).reshape(-1, 3))
    return a, a.shape[0]


def _call(fn, free, xyz1, xyz2, cutoff):
    a, n1 = _prep(xyz1)
    b, n2 = _prep(xyz2)
    row_ptr = np.zeros(n1 + 1, np.int64)
    cols_p = ctypes.c_void_p()
    rc = fn(a.ctypes.data, n1, b.ctypes.data, n2, float(cutoff), row_ptr.ctypes.data, ctypes.byref(cols_p))
    if rc != 0:
        raise RuntimeError("native search failed with code %d" % rc)
    total = int(row_ptr[-1])
    if total and cols_p.value:
        cols = np.ctypeslib.as_array(ctypes.cast(cols_p, ctypes.POINTER(ctypes.c_int32)), shape=(total,)).copy()
    else:
        cols = np.zeros(0, np.int32)
    if cols_p.value:
        free(cols_p)
    return row_ptr, cols


def oracle_find_contacts(xyz1, xyz2, cutoff):
    lib = _load_oracle()
    return _call(lib.oracle_find_contacts, lib.oracle_free, xyz1, xyz2, cutoff)


def ref_find_contacts(xyz1, xyz2, cutoff):
    """The compiled reference.  Raises the soft stack limit first: the reference keeps
    4*(n1+n2) bytes of flags on the stack (gridsearch.C:413-416)."""
    lib = _load_ref()
    n = np.asarray(xyz1).size // 3 + np.asarray(xyz2).size // 3
    if 4 * n > 4 * 1024 * 1024:
        soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
        want = resource.RLIM_INFINITY if hard == resource.RLIM_INFINITY else hard
        if soft != want:
            try:
                resource.setrlimit(resource.RLIMIT_STACK, (want, hard))
            except (ValueError, OSError):
                pass
    return _call(lib.ref_find_contacts_csr, lib.ref_free, xyz1, xyz2, cutoff)


def as_list_of_lists(row_ptr, cols, n2):
    """Shape of cy_find_contacts' return value: n1+n2 rows, last n2 empty (App. A.6 Q3)."""
    out = [cols[row_ptr[i]:row_ptr[i + 1]].tolist() for i in range(len(row_ptr) - 1)]
    out.extend([] for _ in range(n2))
    return out


def grid_geometry(xyz1, xyz2, cutoff):
    """(min[3] f32, inv f32, nb[3] int32) of the reference grid (gridsearch.C:940-985)."""
    lib = _load_oracle()
    a, n1 = _prep(xyz1)
    b, n2 = _prep(xyz2)
    gmin = np.zeros(3, np.float32)
    inv = np.zeros(1, np.float32)
    nb = np.zeros(3, np.int32)
    lib.oracle_grid_geometry(a.ctypes.data, n1, b.ctypes.data, n2, float(cutoff),
                             gmin.ctypes.data, inv.ctypes.data, nb.ctypes.data)
    return gmin, np.float32(inv[0]), nb


def box_coords(xyz, gmin, inv, nb):
    lib = _load_oracle()
    a, n = _prep(xyz)
    out = np.zeros((n, 3), np.int32)
    gmin = np.ascontiguousarray(gmin, np.float32)
    nb = np.ascontiguousarray(nb, np.int32)
    lib.oracle_box_coords(a.ctypes.data, n, gmin.ctypes.data, ctypes.c_float(float(inv)), nb.ctypes.data,
                          out.ctypes.data)
    return out


def stencil_rank_table():
    """rank[dx+1, dy+1, dz+1] of the 27 box offsets in make_neighborlist_sym order."""
    lib = _load_oracle()
    t = np.zeros((3, 3, 3), np.int32)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                t[dx + 1, dy + 1, dz + 1] = lib.oracle_stencil_rank(dx, dy, dz)
    return t


def brute_count(xyz1, xyz2, cutoff):
    lib = _load_oracle()
    a, n1 = _prep(xyz1)
    b, n2 = _prep(xyz2)
    counts = np.zeros(n1, np.int64)
    tot = lib.oracle_brute_count(a.ctypes.data, n1, b.ctypes.data, n2, float(cutoff), counts.ctypes.data)
    return int(tot), counts
