"""ctypes binding of libpycontact_b200.so (C ABI in include/pycontact_b200.h).

The library is the only compute path: if it is missing or cannot be loaded this
module raises -- there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PYCONTACT_B200_LIB: load another build of the same library (tools/phase_clocks.py uses a profiling build)
LIB_PATH = os.environ.get("PYCONTACT_B200_LIB") or os.path.join(_HERE, "libpycontact_b200.so")

PC_OK = 0
PC_ERR_INVALID = -1
PC_ERR_CUDA = -2
PC_ERR_CAPACITY = -3
PC_ERR_MISSING_H = -4
PC_ERR_NOMEM = -5

PC_ATOM_HYDROGEN = 1
PC_ATOM_HBCLASS = 2
PC_ATOM_BACKBONE = 4

PC_LAYOUT_XYZ = 3
PC_LAYOUT_XYZW = 4

ST_HITS_OVERFLOW = 1
ST_HB_OVERFLOW = 2
ST_OUT_OVERFLOW = 4
ST_MISSING_H = 8
ST_BAD_COORDS = 16
ST_TABLE_OVERFLOW = 32
ST_STAGE_OVERFLOW = 64
ST_TAIL_OVERFLOW = 128
ST_PRECISION = 256

# numpy mirrors of the C records
contact_dtype = np.dtype([("i", "<i4"), ("j", "<i4"), ("distance", "<f4"), ("weight", "<f4")])
hbond_dtype = np.dtype([("i", "<i4"), ("j", "<i4"), ("hyd", "<u4"), ("distance", "<f4"), ("angle", "<f4")])
row_dtype = np.dtype([("key", "<u8"), ("score", "<f8"), ("bb1", "<f8"), ("bb2", "<f8"),
                      ("ncontacts", "<u4"), ("nhbonds", "<u4"), ("first", "<u8")])
packed_row_dtype = np.dtype([("key", "<u8"), ("score", "<f4"), ("bb1", "<f4"), ("bb2", "<f4"),
                             ("ncontacts", "<u2"), ("nhbonds", "<u2")])
fold_entry_dtype = np.dtype([("key", "<u8"), ("score", "<f8"), ("bb1", "<f8"), ("bb2", "<f8"), ("hbframes", "<f8")])
slim_row_dtype = np.dtype([("key", "<u4"), ("score", "<f4"), ("ncontacts", "<u2"), ("nhbonds", "<u2")])
assert packed_row_dtype.itemsize == 24 and slim_row_dtype.itemsize == 12
assert contact_dtype.itemsize == 16 and hbond_dtype.itemsize == 20 and row_dtype.itemsize == 48


class PcError(Exception):
    def __init__(self, code, msg):
        super().__init__("pycontact_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/pycontact_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "pc_last_error", "pc_version", "pc_device_count", "pc_create", "pc_destroy", "pc_set_topology", "pc_set_maps", "pc_set_limits", "pc_set_path", "pc_set_fused_accumulate", "pc_set_small_ctas", "pc_set_tail", "pc_set_dense_mode", "pc_set_group_mode", "pc_last_group", "pc_xtc_open", "pc_xtc_info", "pc_xtc_read", "pc_xtc_close",
    "pc_reserve", "pc_scan_device", "pc_accumulate", "pc_scan_host", "pc_batch_totals", "pc_last_status",
    "pc_device_outputs", "pc_fetch", "pc_find_contacts", "pc_free_host", "pc_synth_frames", "pc_host_alloc",
    "pc_host_free", "pc_load_records", "pc_stream_create", "pc_stream_sync", "pc_stream_destroy",
    "pc_device_alloc", "pc_device_free", "pc_memcpy", "pc_memset", "pc_device_sync",
    "pc_launch_count", "pc_event_create", "pc_event_record", "pc_event_elapsed_ms", "pc_event_destroy", "pc_fold_rows", "pc_fetch_packed_rows", "pc_pack_rows", "pc_set_pack_format", "pc_fold_sparse_init", "pc_fold_rows_sparse", "pc_fold_sparse_fetch", "pc_fold_sparse_share", "pc_set_probe", "pc_key_stats", "pc_key_stats_rows", "pc_fold_sparse_export", "pc_fold_sparse_import", "pc_fp32_peak", "pc_mem_info",
    "pc_sasa_points", "pc_sasa", "pc_sasa_device", "pc_find_within", "pc_within_device",
]

_lib = None


def load():
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libpycontact_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python -m pycontact_b200.build`; there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    P = ctypes.POINTER
    lib.pc_last_error.restype = ctypes.c_char_p
    lib.pc_version.restype = ctypes.c_int
    lib.pc_device_count.restype = ctypes.c_int
    lib.pc_create.argtypes = [ctypes.c_int, f64, f64, f64, ctypes.c_int, P(vp)]
    lib.pc_destroy.argtypes = [vp]
    lib.pc_destroy.restype = None
    lib.pc_set_topology.argtypes = [vp, i32, i32] + [vp] * 10
    lib.pc_set_maps.argtypes = [vp, vp, vp, i64]
    lib.pc_set_limits.argtypes = [vp, i32, i32, i32, i32]
    lib.pc_set_path.argtypes = [vp, ctypes.c_int]
    lib.pc_set_fused_accumulate.argtypes = [vp, ctypes.c_int]
    lib.pc_set_small_ctas.argtypes = [vp, ctypes.c_int]
    lib.pc_set_tail.argtypes = [vp, ctypes.c_int]
    lib.pc_set_dense_mode.argtypes = [vp, ctypes.c_int]
    lib.pc_set_group_mode.argtypes = [vp, ctypes.c_int]
    lib.pc_last_group.argtypes = [vp]
    lib.pc_xtc_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.pc_xtc_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i64)]
    lib.pc_xtc_read.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp, vp]
    lib.pc_xtc_close.argtypes = [vp]
    lib.pc_xtc_close.restype = None
    lib.pc_reserve.argtypes = [vp, i32, i64, i64, i64]
    lib.pc_scan_device.argtypes = [vp, vp, vp, i32, i32, i64, i64, i32, vp]
    lib.pc_accumulate.argtypes = [vp, vp]
    lib.pc_scan_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.pc_batch_totals.argtypes = [vp, vp, P(i64), P(i64), P(i64), P(i64)]
    lib.pc_last_status.argtypes = [vp, P(ctypes.c_uint64), P(i64), P(i64), P(i64), P(i32)]
    lib.pc_device_outputs.argtypes = [vp] + [P(vp)] * 10
    lib.pc_fetch.argtypes = [vp, vp] + [vp] * 10
    lib.pc_find_contacts.argtypes = [ctypes.c_int, vp, i32, vp, i32, f64, vp, P(vp)]
    lib.pc_free_host.argtypes = [vp]
    lib.pc_free_host.restype = None
    lib.pc_synth_frames.argtypes = [ctypes.c_int, vp, vp, vp, vp, vp, i32, ctypes.c_float, ctypes.c_uint64,
                                    i32, i32, i32, i64, vp, vp]
    lib.pc_load_records.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.pc_stream_create.argtypes = [ctypes.c_int, P(vp)]
    lib.pc_stream_sync.argtypes = [vp]
    lib.pc_stream_destroy.argtypes = [vp]
    lib.pc_stream_destroy.restype = None
    lib.pc_device_alloc.argtypes = [ctypes.c_int, i64, P(vp)]
    lib.pc_device_free.argtypes = [ctypes.c_int, vp]
    lib.pc_device_free.restype = None
    lib.pc_memcpy.argtypes = [ctypes.c_int, vp, vp, i64, ctypes.c_int, vp]
    lib.pc_memset.argtypes = [ctypes.c_int, vp, ctypes.c_int, i64, vp]
    lib.pc_device_sync.argtypes = [ctypes.c_int]
    lib.pc_launch_count.restype = ctypes.c_int64
    lib.pc_key_stats.argtypes = [ctypes.c_int, vp, vp, i64, i32, f64, f64, vp, vp]
    lib.pc_key_stats_rows.argtypes = [ctypes.c_int, vp, vp, vp, vp, i64, i32, f64, f64, vp, vp]
    lib.pc_event_create.argtypes = [ctypes.c_int, P(vp)]
    lib.pc_event_record.argtypes = [vp, vp]
    lib.pc_event_elapsed_ms.argtypes = [vp, vp, P(ctypes.c_float)]
    lib.pc_event_destroy.argtypes = [vp]
    lib.pc_event_destroy.restype = None
    lib.pc_fetch_packed_rows.argtypes = [vp, vp, vp, vp, vp]
    lib.pc_fold_sparse_init.argtypes = [vp, i64, vp]
    lib.pc_fold_rows_sparse.argtypes = [vp, vp]
    lib.pc_fold_sparse_share.argtypes = [vp, vp]
    lib.pc_fold_sparse_fetch.argtypes = [vp, vp, vp, i64, P(i64)]
    lib.pc_pack_rows.argtypes = [vp, vp]
    lib.pc_set_pack_format.argtypes = [vp, ctypes.c_int]
    lib.pc_set_probe.argtypes = [vp, vp, vp]
    lib.pc_fold_sparse_export.argtypes = [vp, vp, vp, ctypes.c_int64]
    lib.pc_fold_sparse_import.argtypes = [vp, vp, vp, ctypes.c_int64]
    lib.pc_fp32_peak.argtypes = [ctypes.c_int, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
    lib.pc_mem_info.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    lib.pc_fold_rows.argtypes = [vp, vp, i64, vp]
    f32 = ctypes.c_float
    lib.pc_sasa_points.argtypes = [i32, i32, vp]
    lib.pc_sasa.argtypes = [ctypes.c_int, vp, i32, i32, vp, f32, i32, f64, i32, i32, vp, vp, vp]
    lib.pc_sasa_device.argtypes = [ctypes.c_int, vp, i32, i32, i64, vp, f32, i32, f64, i32, i32, vp, vp, vp, vp, vp]
    lib.pc_find_within.argtypes = [ctypes.c_int, vp, i32, i32, vp, vp, f32, vp]
    lib.pc_within_device.argtypes = [ctypes.c_int, vp, i32, i32, i64, vp, vp, f32, vp, vp]
    lib.pc_host_alloc.argtypes = [P(vp), i64]
    lib.pc_host_free.argtypes = [vp]
    lib.pc_host_free.restype = None
    _lib = lib
    return lib


def check(rc):
    if rc != PC_OK:
        raise PcError(rc, load().pc_last_error().decode("utf-8", "replace"))


def ptr(a):
    """host pointer of a numpy array (None -> NULL)"""
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class PinnedArray:
    """A numpy array over cudaHostAlloc'd memory (freed with the object)."""

    def __init__(self, shape, dtype):
        lib = load()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = ctypes.c_void_p()
        check(lib.pc_host_alloc(ctypes.byref(p), max(nbytes, 16)))
        self._p = p
        buf = (ctypes.c_byte * max(nbytes, 16)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        try:
            if self._p:
                self.array = None
                load().pc_host_free(self._p)
                self._p = None
        except Exception:
            pass


class DeviceBuffer:
    """cudaMalloc'd memory owned by a Python object."""

    def __init__(self, nbytes, device=0):
        self.device = int(device)
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(load().pc_device_alloc(self.device, max(self.nbytes, 16), ctypes.byref(p)))
        self.ptr = p

    @classmethod
    def from_array(cls, arr, device=0):
        arr = np.ascontiguousarray(arr)
        buf = cls(arr.nbytes, device)
        buf.upload(arr)
        return buf

    def upload(self, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        check(load().pc_memcpy(self.device, ctypes.c_void_p(self.ptr.value + offset), ptr(arr), arr.nbytes, 1, None))
        check(load().pc_device_sync(self.device))

    def download(self, dtype, count, offset=0):
        out = np.empty(int(count), dtype)
        check(load().pc_memcpy(self.device, ptr(out), ctypes.c_void_p(self.ptr.value + offset), out.nbytes, 2, None))
        check(load().pc_device_sync(self.device))
        return out

    def at(self, offset):
        return ctypes.c_void_p(self.ptr.value + int(offset))

    def free(self):
        if self.ptr:
            load().pc_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
