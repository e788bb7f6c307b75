"""ctypes driver for oracle/_ref/libnddwt_ref.so -- the reference's own mex/nddwt.c compiled here
(oracle/Makefile).  TEST INFRASTRUCTURE ONLY (tests, smoke, bench CPU-baseline legs).

It plays the role of the MATLAB side of the 'mex' compute path: builds the scale2-scaled stored
filters as `get_filters` does, takes fftn(x) as `dec` does (nd_dwt_2D.m:156), marshals split
real/imag column-major doubles and int dims as mexFunction does (mex/nd_dwt_mex.c:55-76,90-98,141-148),
and calls nd_dwt_dec / nd_dwt_rec / *_1level.
"""
from __future__ import annotations

import ctypes
import os
import numpy as np

from . import nddwt_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libnddwt_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.nd_dwt_dec_1level.argtypes = [dp] * 6 + [ctypes.c_int, ip]
        L.nd_dwt_rec_1level.argtypes = [dp] * 6 + [ctypes.c_int, ip, ctypes.c_int]
        L.nd_dwt_dec.argtypes = [dp] * 6 + [ctypes.c_int, ip, ctypes.c_int]
        L.nd_dwt_rec.argtypes = [dp] * 6 + [ctypes.c_int, ip, ctypes.c_int, ctypes.c_int]
        for f in (L.nd_dwt_dec_1level, L.nd_dwt_rec_1level, L.nd_dwt_dec, L.nd_dwt_rec):
            f.restype = None
        _lib = L
    return _lib


def _split(a):
    """complex ndarray (MATLAB shape) -> column-major split real / imag double buffers."""
    a = np.asarray(a, dtype=np.complex128)
    re = np.ascontiguousarray(a.real.ravel(order="F"))
    im = np.ascontiguousarray(a.imag.ravel(order="F"))
    return re, im


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def mex_call(x_f, f_dec, direction, level, pres_l2):
    """y = nd_dwt_mex(x_f, f_dec, dir, level, pres_l2) on the compiled reference core."""
    L = lib()
    fr, fi = _split(f_dec)
    xr, xi = _split(x_f)
    if direction == 0:
        sizes = x_f.shape
        d = len(sizes)
        nb = orc.num_bands(d, level)
        numel = int(np.prod(sizes))
        outr = np.zeros(numel * nb)
        outi = np.zeros(numel * nb)
        dims = (ctypes.c_int * d)(*sizes)
        if level == 1:
            L.nd_dwt_dec_1level(_ptr(outr), _ptr(outi), _ptr(xr), _ptr(xi), _ptr(fr), _ptr(fi), d, dims)
        else:
            L.nd_dwt_dec(_ptr(outr), _ptr(outi), _ptr(xr), _ptr(xi), _ptr(fr), _ptr(fi), d, dims, level)
        shape = tuple(sizes) + (nb,)
    else:
        sizes = x_f.shape[:-1]
        d = len(sizes)
        numel = int(np.prod(sizes))
        outr = np.zeros(numel)
        outi = np.zeros(numel)
        dims = (ctypes.c_int * d)(*sizes)
        if level == 1:
            L.nd_dwt_rec_1level(_ptr(outr), _ptr(outi), _ptr(xr), _ptr(xi), _ptr(fr), _ptr(fi), d, dims, int(pres_l2))
        else:
            L.nd_dwt_rec(_ptr(outr), _ptr(outi), _ptr(xr), _ptr(xi), _ptr(fr), _ptr(fi), d, dims, level, int(pres_l2))
        shape = tuple(sizes)
    return (outr + 1j * outi).reshape(shape, order="F")


def dec(x, wname, level, pres_l2_norm=False):
    """obj.dec(x, level) with compute='mex' (nd_dwt_2D.m:156-160), on the compiled reference core."""
    x = np.asarray(x)
    d = x.ndim
    f = orc.get_filters(wname, x.shape, pres_l2_norm, mex_scale=True)
    x_f = orc._fftn(x.astype(np.complex128), tuple(range(d)))
    y = mex_call(x_f, f, 0, level, pres_l2_norm)
    return y.real if not np.iscomplexobj(x) else y


def rec(y, wname, pres_l2_norm=False):
    """obj.rec(y) with compute='mex' (nd_dwt_2D.m:215-222)."""
    y = np.asarray(y)
    d = y.ndim - 1
    level = orc.infer_level(d, y.shape[-1])
    f = orc.get_filters(wname, y.shape[:-1], pres_l2_norm, mex_scale=True)
    c_f = orc._fftn(y.astype(np.complex128), tuple(range(d)))
    out = mex_call(c_f, f, 1, level, pres_l2_norm)
    return out.real if not np.iscomplexobj(y) else out
