"""ctypes binding of libnddwt_b200.so (the C ABI in include/nddwt_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `csrc/Makefile`.  There is no
fallback: if the shared object is missing, importing this module's `lib()` raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnddwt_b200.so")

NDDWT_F32, NDDWT_F64, NDDWT_C64, NDDWT_C128 = 0, 1, 2, 3
ERR_ARG, ERR_WAVELET, ERR_SHORT_DIM, ERR_CUDA, ERR_NOMEM, ERR_SIZE = -1, -2, -3, -4, -5, -6

# every symbol include/nddwt_b200.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "nddwt_last_error", "nddwt_version", "nddwt_wave_filters", "nddwt_num_bands", "nddwt_infer_level",
    "nddwt_plan_create", "nddwt_plan_create_slab", "nddwt_plan_destroy", "nddwt_plan_set_dilations",
    "nddwt_plan_set_batch", "nddwt_plan_set_kernel_mode", "nddwt_plan_set_param", "nddwt_plan_set_shrink",
    "nddwt_mplan_set_shrink", "nddwt_plan_launch_count", "nddwt_plan_last_path",
    "nddwt_plan_profile", "nddwt_plan_kernel_time", "nddwt_plan_last_synthesis_kernel",
    "nddwt_dec", "nddwt_rec", "nddwt_shrink", "nddwt_dec_host", "nddwt_rec_host",
    "nddwt_mplan_create", "nddwt_mplan_create_rank", "nddwt_mplan_export_size", "nddwt_mplan_export",
    "nddwt_mplan_import", "nddwt_mplan_destroy", "nddwt_mplan_world", "nddwt_mplan_num_local", "nddwt_mplan_slab",
    "nddwt_mplan_is_separable", "nddwt_mplan_set_dilations", "nddwt_mplan_set_kernel_mode", "nddwt_mplan_set_param",
    "nddwt_mplan_dec", "nddwt_mplan_rec", "nddwt_mplan_sync", "nddwt_mplan_launch_count", "nddwt_mplan_halo_bytes",
    "nddwt_mplan_wait_timeouts", "nddwt_slab_route", "nddwt_mplan_dec_host", "nddwt_mplan_rec_host", "nddwt_mplan_profile", "nddwt_mplan_kernel_time",
    "nddwt_halo_planes", "nddwt_dec_level_slab", "nddwt_plan_is_separable", "nddwt_dec_level_slab_part", "nddwt_rec_level_slab_stage1_part", "nddwt_rec_level_slab_stage2_scatter", "nddwt_accumulate", "nddwt_rec_level_slab_stage1", "nddwt_rec_level_slab_stage2",
]

_lib = None


class NddwtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libnddwt_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C non-decimated_wavelets_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i64p, ip, dp = c.c_void_p, c.POINTER(c.c_int64), c.POINTER(c.c_int), c.POINTER(c.c_double)
    L.nddwt_last_error.restype = c.c_char_p
    L.nddwt_version.restype = c.c_char_p
    L.nddwt_wave_filters.argtypes = [c.c_char_p, dp, dp, ip]
    L.nddwt_num_bands.argtypes = [c.c_int, c.c_int]
    L.nddwt_num_bands.restype = c.c_int64
    L.nddwt_infer_level.argtypes = [c.c_int, c.c_int64]
    L.nddwt_plan_create.argtypes = [c.POINTER(vp), c.c_int, i64p, c.POINTER(c.c_char_p), c.c_int, c.c_int, c.c_int]
    L.nddwt_plan_create_slab.argtypes = [c.POINTER(vp), c.c_int, i64p, c.c_int64, c.POINTER(c.c_char_p), c.c_int,
                                         c.c_int, c.c_int]
    L.nddwt_plan_destroy.argtypes = [vp]
    L.nddwt_plan_set_dilations.argtypes = [vp, ip, c.c_int]
    L.nddwt_plan_set_batch.argtypes = [vp, c.c_int64]
    L.nddwt_plan_set_kernel_mode.argtypes = [vp, c.c_int]
    L.nddwt_plan_set_param.argtypes = [vp, c.c_char_p, c.c_int64]
    L.nddwt_plan_set_shrink.argtypes = [vp, c.c_int, dp, c.c_int]
    L.nddwt_mplan_set_shrink.argtypes = [vp, c.c_int, dp, c.c_int]
    L.nddwt_plan_launch_count.argtypes = [vp]
    L.nddwt_plan_launch_count.restype = c.c_int64
    L.nddwt_plan_last_path.argtypes = [vp]
    L.nddwt_plan_last_synthesis_kernel.argtypes = [vp]
    L.nddwt_plan_profile.argtypes = [vp, c.c_int]
    L.nddwt_plan_kernel_time.argtypes = [vp, c.c_int, dp, c.POINTER(c.c_int64)]
    L.nddwt_dec.argtypes = [vp, vp, vp, c.c_int, vp]
    L.nddwt_rec.argtypes = [vp, vp, vp, c.c_int, vp]
    L.nddwt_shrink.argtypes = [vp, vp, c.c_int, vp]
    L.nddwt_dec_host.argtypes = [vp, vp, vp, c.c_int]
    L.nddwt_rec_host.argtypes = [vp, vp, vp, c.c_int]
    L.nddwt_halo_planes.argtypes = [vp, c.c_int, ip, ip]
    L.nddwt_dec_level_slab.argtypes = [vp, c.c_int, vp, vp, vp, c.POINTER(vp), vp]
    L.nddwt_plan_is_separable.argtypes = [vp]
    L.nddwt_dec_level_slab_part.argtypes = [vp, c.c_int, c.c_int, vp, vp, vp, c.POINTER(vp), vp]
    L.nddwt_rec_level_slab_stage1_part.argtypes = [vp, c.c_int, c.c_int, c.POINTER(vp), vp, vp, vp]
    L.nddwt_rec_level_slab_stage2_scatter.argtypes = [vp, c.c_int, vp, vp, vp, vp, vp, vp]
    L.nddwt_accumulate.argtypes = [vp, vp, vp, c.c_int64, vp]
    L.nddwt_rec_level_slab_stage1.argtypes = [vp, c.c_int, c.POINTER(vp), vp, vp, vp]
    L.nddwt_rec_level_slab_stage2.argtypes = [vp, c.c_int, vp, vp, vp, vp, vp, vp]
    L.nddwt_mplan_create.argtypes = [c.POINTER(vp), c.c_int, i64p, c.POINTER(c.c_char_p), c.c_int, c.c_int, c.c_int, ip]
    L.nddwt_mplan_create_rank.argtypes = [c.POINTER(vp), c.c_int, i64p, c.POINTER(c.c_char_p), c.c_int, c.c_int,
                                          c.c_int, c.c_int, c.c_int]
    L.nddwt_mplan_export_size.restype = c.c_int64
    L.nddwt_mplan_export.argtypes = [vp, vp]
    L.nddwt_mplan_import.argtypes = [vp, vp]
    L.nddwt_mplan_destroy.argtypes = [vp]
    L.nddwt_mplan_world.argtypes = [vp]
    L.nddwt_mplan_num_local.argtypes = [vp]
    L.nddwt_mplan_slab.argtypes = [vp, c.c_int, i64p, i64p]
    L.nddwt_mplan_is_separable.argtypes = [vp]
    L.nddwt_mplan_set_dilations.argtypes = [vp, ip, c.c_int]
    L.nddwt_mplan_set_kernel_mode.argtypes = [vp, c.c_int]
    L.nddwt_mplan_set_param.argtypes = [vp, c.c_char_p, c.c_int64]
    L.nddwt_mplan_dec.argtypes = [vp, c.POINTER(vp), c.POINTER(vp), c.c_int, c.POINTER(vp)]
    L.nddwt_mplan_rec.argtypes = [vp, c.POINTER(vp), c.POINTER(vp), c.c_int, c.POINTER(vp)]
    L.nddwt_mplan_sync.argtypes = [vp]
    L.nddwt_mplan_dec_host.argtypes = [vp, vp, vp, c.c_int]
    L.nddwt_mplan_rec_host.argtypes = [vp, vp, vp, c.c_int]
    L.nddwt_mplan_launch_count.argtypes = [vp]
    L.nddwt_mplan_launch_count.restype = c.c_int64
    L.nddwt_mplan_halo_bytes.argtypes = [vp]
    L.nddwt_mplan_halo_bytes.restype = c.c_int64
    L.nddwt_mplan_wait_timeouts.argtypes = [vp]
    L.nddwt_mplan_profile.argtypes = [vp, c.c_int]
    L.nddwt_mplan_kernel_time.argtypes = [vp, c.c_int, c.c_int, dp, c.POINTER(c.c_int64)]
    L.nddwt_slab_route.argtypes = [c.c_int64, c.c_int, c.c_int, c.c_int, c.c_int64, c.c_int64, i64p, c.c_int]
    _lib = L
    return L


def _set_shrink(fn, handle, table, ndims):
    import numpy as np
    if table is None:
        check(fn(handle, 0, None, 0))
        return
    t = np.ascontiguousarray(np.asarray(table, dtype=np.float64))
    if t.ndim != 2 or t.shape[1] != (1 << ndims):
        raise ValueError("threshold table must be [nlevels][2^ndims]")
    check(fn(handle, 1, t.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), t.shape[0]))


def slab_route(n_last, world, rank, which, below, above):
    """Host-only: runs (owner, first local plane, count, first halo slot) of rank's halo (nddwt_slab_route)."""
    cap = 4 * (int(below) + int(above) + 1)
    buf = (ctypes.c_int64 * (4 * cap))()
    n = lib().nddwt_slab_route(int(n_last), int(world), int(rank), int(which), int(below), int(above), buf, cap)
    if n < 0:
        check(n)
    return [tuple(int(buf[4 * k + i]) for i in range(4)) for k in range(n)]


def check(rc):
    if rc != 0:
        raise NddwtError(rc, lib().nddwt_last_error().decode("utf-8", "replace"))


def wave_filters_c(wname: str):
    import numpy as np
    lo = (ctypes.c_double * 20)()
    hi = (ctypes.c_double * 20)()
    n = ctypes.c_int(0)
    check(lib().nddwt_wave_filters(str(wname).encode(), lo, hi, ctypes.byref(n)))
    return np.array(lo[: n.value]), np.array(hi[: n.value])


class Plan:
    """Owns one nddwt_plan handle (the stored-filter object, reused across dec/rec calls)."""

    def __init__(self, dims, wnames, dtype_code, pres_l2_norm, device=0, global_last=None):
        L = lib()
        self.ndims = len(dims)
        self.dims = tuple(int(d) for d in dims)
        arr = (ctypes.c_int64 * self.ndims)(*self.dims)
        names = (ctypes.c_char_p * self.ndims)(*[str(w).encode() for w in wnames])
        h = ctypes.c_void_p()
        if global_last is None:
            rc = L.nddwt_plan_create(ctypes.byref(h), self.ndims, arr, names, dtype_code, int(bool(pres_l2_norm)), device)
        else:
            rc = L.nddwt_plan_create_slab(ctypes.byref(h), self.ndims, arr, int(global_last), names, dtype_code,
                                          int(bool(pres_l2_norm)), device)
        check(rc)
        self.handle = h
        self.dtype_code = dtype_code
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            lib().nddwt_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_dilations(self, dil):
        a = (ctypes.c_int * len(dil))(*[int(v) for v in dil])
        check(lib().nddwt_plan_set_dilations(self.handle, a, len(dil)))

    def set_batch(self, batch):
        check(lib().nddwt_plan_set_batch(self.handle, int(batch)))

    def set_kernel_mode(self, mode):
        check(lib().nddwt_plan_set_kernel_mode(self.handle, int(mode)))

    def set_param(self, name, value):
        check(lib().nddwt_plan_set_param(self.handle, str(name).encode(), int(value)))

    def set_shrink(self, table):
        """table: None (off) or [nlevels][2^ndims] soft thresholds (level 1 = finest first; column 0 ignored)."""
        _set_shrink(lib().nddwt_plan_set_shrink, self.handle, table, self.ndims)

    @property
    def launches(self):
        return int(lib().nddwt_plan_launch_count(self.handle))

    @property
    def separable(self):
        return bool(lib().nddwt_plan_is_separable(self.handle))

    def profile(self, on):
        check(lib().nddwt_plan_profile(self.handle, int(bool(on))))

    def kernel_time(self, kind):
        """(total_ms, launches) of one kernel kind since the last read (0 dec3, 1 rec3, 2 dec_last, 3 rec_last, 4 generic)."""
        ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
        check(lib().nddwt_plan_kernel_time(self.handle, kind, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    @property
    def last_path(self):
        return int(lib().nddwt_plan_last_path(self.handle))

    @property
    def last_synthesis_kernel(self):
        """0 none, 1 direct-load tiles, 2 TMA-staged 32-column tiles, 4 full-row tiles."""
        return int(lib().nddwt_plan_last_synthesis_kernel(self.handle))

    def dec(self, x_ptr, c_ptr, level, stream=0):
        check(lib().nddwt_dec(self.handle, x_ptr, c_ptr, level, stream))

    def rec(self, c_ptr, x_ptr, level, stream=0):
        check(lib().nddwt_rec(self.handle, c_ptr, x_ptr, level, stream))

    def shrink(self, c_ptr, level, stream=0):
        check(lib().nddwt_shrink(self.handle, c_ptr, level, stream))

    def dec_host(self, x_ptr, c_ptr, level):
        check(lib().nddwt_dec_host(self.handle, x_ptr, c_ptr, level))

    def rec_host(self, c_ptr, x_ptr, level):
        check(lib().nddwt_rec_host(self.handle, c_ptr, x_ptr, level))

    def halo_planes(self, level_index):
        b, a = ctypes.c_int(0), ctypes.c_int(0)
        check(lib().nddwt_halo_planes(self.handle, level_index, ctypes.byref(b), ctypes.byref(a)))
        return b.value, a.value

    def dec_level_slab(self, level_index, a_in, halo_lo, halo_hi, out_ptrs, stream=0):
        arr = (ctypes.c_void_p * len(out_ptrs))(*out_ptrs)
        check(lib().nddwt_dec_level_slab(self.handle, level_index, a_in, halo_lo, halo_hi, arr, stream))

    def dec_level_slab_part(self, level_index, part, a_in, halo_lo, halo_hi, out_ptrs, stream=0):
        arr = (ctypes.c_void_p * len(out_ptrs))(*out_ptrs)
        check(lib().nddwt_dec_level_slab_part(self.handle, level_index, part, a_in, halo_lo, halo_hi, arr, stream))

    def rec_level_slab_stage1_part(self, level_index, part, in_ptrs, u_lo, u_hi, stream=0):
        arr = (ctypes.c_void_p * len(in_ptrs))(*in_ptrs)
        check(lib().nddwt_rec_level_slab_stage1_part(self.handle, level_index, part, arr, u_lo, u_hi, stream))

    def rec_level_slab_stage2_scatter(self, level_index, u_lo, u_hi, a_out, over_lo, over_hi, stream=0):
        check(lib().nddwt_rec_level_slab_stage2_scatter(self.handle, level_index, u_lo, u_hi, a_out, over_lo, over_hi, stream))

    def accumulate(self, dst, src, nelem, stream=0):
        check(lib().nddwt_accumulate(self.handle, dst, src, nelem, stream))

    def rec_level_slab_stage1(self, level_index, in_ptrs, u_lo, u_hi, stream=0):
        arr = (ctypes.c_void_p * len(in_ptrs))(*in_ptrs)
        check(lib().nddwt_rec_level_slab_stage1(self.handle, level_index, arr, u_lo, u_hi, stream))

    def rec_level_slab_stage2(self, level_index, u_lo, u_hi, halo_lo, halo_hi, a_out, stream=0):
        check(lib().nddwt_rec_level_slab_stage2(self.handle, level_index, u_lo, u_hi, halo_lo, halo_hi, a_out, stream))


class MultiPlan:
    """Owns one nddwt_mplan handle: the multi-GPU plan (slabs along the last dim, peer-memory halo pushes).

    MultiPlan(dims, wnames, dtype, l2, devices=[...])            one process drives the listed devices
    MultiPlan(dims, wnames, dtype, l2, rank=r, world=P, device=d) one rank of a one-process-per-GPU job; call
        `connect(all_gather)` before the first transform (all_gather: bytes -> list of every rank's bytes)."""

    def __init__(self, dims, wnames, dtype_code, pres_l2_norm, devices=None, rank=None, world=None, device=0):
        L = lib()
        self.ndims = len(dims)
        self.dims = tuple(int(d) for d in dims)
        arr = (ctypes.c_int64 * self.ndims)(*self.dims)
        names = (ctypes.c_char_p * self.ndims)(*[str(w).encode() for w in wnames])
        h = ctypes.c_void_p()
        if rank is None:
            devices = list(devices if devices is not None else [0])
            darr = (ctypes.c_int * len(devices))(*devices)
            check(L.nddwt_mplan_create(ctypes.byref(h), self.ndims, arr, names, dtype_code, int(bool(pres_l2_norm)),
                                       len(devices), darr))
            self.devices = devices
        else:
            check(L.nddwt_mplan_create_rank(ctypes.byref(h), self.ndims, arr, names, dtype_code,
                                            int(bool(pres_l2_norm)), int(rank), int(world), int(device)))
            self.devices = [device]
        self.handle = h
        self.dtype_code = dtype_code
        self.rank = rank
        self.world = int(L.nddwt_mplan_world(h))
        self.num_local = int(L.nddwt_mplan_num_local(h))

    def close(self):
        if getattr(self, "handle", None):
            lib().nddwt_mplan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def connect(self, all_gather):
        """Exchange the IPC export blobs (rank plans).  all_gather(bytes) -> [bytes of rank 0, ..., rank P-1]."""
        n = int(lib().nddwt_mplan_export_size())
        blob = (ctypes.c_char * n)()
        check(lib().nddwt_mplan_export(self.handle, blob))
        blobs = all_gather(bytes(blob))
        joined = b"".join(blobs)
        assert len(joined) == n * self.world
        buf = (ctypes.c_char * len(joined)).from_buffer_copy(joined)
        check(lib().nddwt_mplan_import(self.handle, buf))

    def slab(self, rank):
        s, c = ctypes.c_int64(0), ctypes.c_int64(0)
        check(lib().nddwt_mplan_slab(self.handle, rank, ctypes.byref(s), ctypes.byref(c)))
        return s.value, c.value

    @property
    def separable(self):
        return bool(lib().nddwt_mplan_is_separable(self.handle))

    def set_dilations(self, dil):
        a = (ctypes.c_int * len(dil))(*[int(v) for v in dil])
        check(lib().nddwt_mplan_set_dilations(self.handle, a, len(dil)))

    def set_kernel_mode(self, mode):
        check(lib().nddwt_mplan_set_kernel_mode(self.handle, int(mode)))

    def set_shrink(self, table):
        _set_shrink(lib().nddwt_mplan_set_shrink, self.handle, table, self.ndims)

    def set_param(self, name, value):
        check(lib().nddwt_mplan_set_param(self.handle, str(name).encode(), int(value)))

    @staticmethod
    def _ptrs(vals):
        return (ctypes.c_void_p * len(vals))(*[int(v) if v else None for v in vals])

    def dec(self, x_ptrs, c_ptrs, level, streams=None):
        st = self._ptrs(streams) if streams is not None else None
        check(lib().nddwt_mplan_dec(self.handle, self._ptrs(x_ptrs), self._ptrs(c_ptrs), int(level), st))

    def rec(self, c_ptrs, x_ptrs, level, streams=None):
        st = self._ptrs(streams) if streams is not None else None
        check(lib().nddwt_mplan_rec(self.handle, self._ptrs(c_ptrs), self._ptrs(x_ptrs), int(level), st))

    def sync(self):
        check(lib().nddwt_mplan_sync(self.handle))

    def dec_host(self, x_ptr, c_ptr, level):
        check(lib().nddwt_mplan_dec_host(self.handle, x_ptr, c_ptr, int(level)))

    def rec_host(self, c_ptr, x_ptr, level):
        check(lib().nddwt_mplan_rec_host(self.handle, c_ptr, x_ptr, int(level)))

    @property
    def launches(self):
        return int(lib().nddwt_mplan_launch_count(self.handle))

    def profile(self, on):
        check(lib().nddwt_mplan_profile(self.handle, int(bool(on))))

    def kernel_time(self, kind, local_index=0):
        ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
        check(lib().nddwt_mplan_kernel_time(self.handle, local_index, kind, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    @property
    def halo_bytes(self):
        return int(lib().nddwt_mplan_halo_bytes(self.handle))

    @property
    def wait_timeouts(self):
        return int(lib().nddwt_mplan_wait_timeouts(self.handle))
