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
    "nddwt_plan_set_batch", "nddwt_plan_set_kernel_mode", "nddwt_plan_set_param", "nddwt_plan_launch_count", "nddwt_plan_last_path",
    "nddwt_plan_profile", "nddwt_plan_kernel_time", "nddwt_plan_last_synthesis_kernel",
    "nddwt_dec", "nddwt_rec", "nddwt_dec_host", "nddwt_rec_host",
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
    L.nddwt_plan_launch_count.argtypes = [vp]
    L.nddwt_plan_launch_count.restype = c.c_int64
    L.nddwt_plan_last_path.argtypes = [vp]
    L.nddwt_plan_last_synthesis_kernel.argtypes = [vp]
    L.nddwt_plan_profile.argtypes = [vp, c.c_int]
    L.nddwt_plan_kernel_time.argtypes = [vp, c.c_int, dp, c.POINTER(c.c_int64)]
    L.nddwt_dec.argtypes = [vp, vp, vp, c.c_int, vp]
    L.nddwt_rec.argtypes = [vp, vp, vp, c.c_int, vp]
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
    _lib = L
    return L


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
