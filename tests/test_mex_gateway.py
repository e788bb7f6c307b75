"""EXECUTES the MEX gateway (non-decimated_wavelets_b200/matlab/nd_dwt_mex.cpp -- what a MATLAB user builds with
mex / mexcuda, the drop-in for the reference's mex/nd_dwt_mex.c) against a mock MEX runtime (tests/mexmock/mock_mex.cpp:
the C Matrix API subset the gateway uses, mexErrMsgIdAndTxt that does not return, and the mxGPU API backed by real
device memory).  MATLAB itself is not available here; this is the closest executable check of the MATLAB-facing
boundary: argument handling, plan handles, the descriptor-struct compatibility path, output shapes, the error texts
of the reference (mex/nd_dwt_mex.c:19-51,124-127), the three build flavours (interleaved complex, legacy split
complex, + gpuArray branch), and the results against the oracle.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import nddwt_b200 as nd
from oracle import nddwt_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "non-decimated_wavelets_b200")
MOCK = os.path.join(ROOT, "tests", "mexmock")
SRCS = [os.path.join(PKG, "matlab", "nd_dwt_mex.cpp"), os.path.join(MOCK, "mock_mex.cpp")]
CLS_OF = {np.dtype(np.float64): 6, np.dtype(np.complex128): 6, np.dtype(np.float32): 7, np.dtype(np.complex64): 7,
          np.dtype(np.uint64): 15}
REAL_OF = {6: np.float64, 7: np.float32, 15: np.uint64}
FLAGS = {"interleaved": [], "split": ["-DNDDWT_MEX_SPLIT_COMPLEX"],
         "gpu": ["-DNDDWT_MEX_GPU", "-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"]}


def _build(flavour):
    """g++ of the gateway + the mock runtime into one shared library (kept in-tree so that it travels to the GPU box)."""
    out_dir = os.path.join(MOCK, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libmexmock_%s.so" % flavour)
    deps = SRCS + [os.path.join(PKG, "matlab", "stub", "mex.h"), os.path.join(ROOT, "include", "nddwt_b200.h"), nd.LIB_PATH]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    cmd = ["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-Wall", "-Wextra", "-I", os.path.join(PKG, "matlab", "stub"),
           "-o", so] + SRCS + ["-L", PKG, "-lnddwt_b200", "-Wl,-rpath,$ORIGIN/../../../non-decimated_wavelets_b200"] + FLAGS[flavour]
    try:
        subprocess.check_call(cmd)
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(so):          # a prebuilt library that merely looks stale (copied tree) is still usable
            raise
    return so


class MexError(Exception):
    def __init__(self, ident, msg):
        super().__init__(msg)
        self.ident, self.msg = ident, msg


class Mex:
    """ctypes driver of one build flavour: numpy <-> mxArray, y = mex(x, f, dir, level, l2)."""

    def __init__(self, flavour):
        self.flavour = flavour
        L = self.L = ctypes.CDLL(_build(flavour))
        vp, ci, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
        for name, res, args in [("mock_create_numeric", vp, [ci, ctypes.POINTER(cu), ci, ci]), ("mock_create_string", vp, [ctypes.c_char_p]),
                                ("mock_create_cell", vp, [ci]), ("mock_set_cell", None, [vp, ci, vp]), ("mock_create_struct", vp, []),
                                ("mock_set_field", None, [vp, ctypes.c_char_p, vp]), ("mock_real", vp, [vp]), ("mock_imag", vp, [vp]),
                                ("mock_class", ci, [vp]), ("mock_is_complex", ci, [vp]), ("mock_is_gpu", ci, [vp]), ("mock_ndims", ci, [vp]),
                                ("mock_dim", cu, [vp, ci]), ("mock_split_complex", ci, []), ("mock_live_mallocs", ci, []),
                                ("mock_destroy", None, [vp]), ("mock_run_atexit", None, []),
                                ("mock_call", ci, [ci, ctypes.POINTER(vp), ci, ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_char_p, ci])]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if flavour == "gpu":
            for name, res, args in [("mock_gpu_array", vp, [vp]), ("mock_gather", vp, [vp]), ("mock_live_gpu_handles", ci, [])]:
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        self.split = bool(L.mock_split_complex())

    # ---- numpy / python -> mxArray
    def to_mx(self, v):
        L = self.L
        if isinstance(v, str):
            return L.mock_create_string(v.encode())
        if isinstance(v, dict):                       # 1 x 1 struct
            s = L.mock_create_struct()
            for k, val in v.items():
                L.mock_set_field(s, k.encode(), self.to_mx(val))
            return s
        if isinstance(v, (list, tuple)) and v and isinstance(v[0], str):      # cellstr
            c = L.mock_create_cell(len(v))
            for i, sv in enumerate(v):
                L.mock_set_cell(c, i, self.to_mx(sv))
            return c
        a = np.asarray(v, dtype=np.float64) if not isinstance(v, np.ndarray) else v
        dims = list(a.shape) if a.ndim >= 2 else ([a.shape[0], 1] if a.ndim == 1 else [1, 1])
        if a.size == 0:
            dims = [0, 0]
        cplx = np.iscomplexobj(a)
        cls = CLS_OF[a.dtype]
        m = L.mock_create_numeric(len(dims), (ctypes.c_uint64 * len(dims))(*dims), cls, int(cplx))
        if a.size:
            flat = np.asfortranarray(a).ravel(order="F")
            rt = REAL_OF[cls]
            if cplx and self.split:
                re, im = np.ascontiguousarray(flat.real, dtype=rt), np.ascontiguousarray(flat.imag, dtype=rt)   # (kept alive)
                ctypes.memmove(L.mock_real(m), re.ctypes.data, re.nbytes)
                ctypes.memmove(L.mock_imag(m), im.ctypes.data, im.nbytes)
            else:
                ctypes.memmove(L.mock_real(m), flat.ctypes.data, flat.nbytes)
        return m

    def from_mx(self, m):
        L = self.L
        dims = [int(L.mock_dim(m, i)) for i in range(L.mock_ndims(m))]
        rt = REAL_OF[L.mock_class(m)]
        n = int(np.prod(dims))
        cplx = bool(L.mock_is_complex(m))
        if cplx and self.split:
            re = np.ctypeslib.as_array(ctypes.cast(L.mock_real(m), ctypes.POINTER(np.ctypeslib.as_ctypes_type(rt))), (n,)).copy()
            im = np.ctypeslib.as_array(ctypes.cast(L.mock_imag(m), ctypes.POINTER(np.ctypeslib.as_ctypes_type(rt))), (n,)).copy()
            flat = re + 1j * im
            flat = flat.astype(np.complex64 if rt is np.float32 else np.complex128)
        else:
            cnt = n * (2 if cplx else 1)
            raw = np.ctypeslib.as_array(ctypes.cast(L.mock_real(m), ctypes.POINTER(np.ctypeslib.as_ctypes_type(rt))), (cnt,)).copy()
            flat = raw.view(np.complex64 if rt is np.float32 else np.complex128) if cplx else raw
        return flat.reshape(dims, order="F")

    def call(self, *args, nlhs=1, raw_out=False):
        """y = nd_dwt_mex(args...).  numpy / str / dict / list arguments are converted (and destroyed afterwards);
        integers that are mxArray pointers wrapped in `Ptr` are passed through."""
        L = self.L
        mx, own = [], []
        for a in args:
            if isinstance(a, Ptr):
                mx.append(a.p)
            else:
                m = self.to_mx(a)
                mx.append(m)
                own.append(m)
        prhs = (ctypes.c_void_p * max(1, len(mx)))(*mx)
        plhs = (ctypes.c_void_p * max(1, nlhs, 2))()
        eid, emsg = ctypes.create_string_buffer(256), ctypes.create_string_buffer(1024)
        rc = L.mock_call(nlhs, plhs, len(mx), prhs, eid, emsg, 1024)
        for m in own:
            L.mock_destroy(m)
        if rc:
            raise MexError(eid.value.decode(), emsg.value.decode())
        outs = []
        for i in range(nlhs):
            if not plhs[i]:
                outs.append(None)
            elif raw_out:
                outs.append(Ptr(plhs[i]))
            else:
                outs.append(self.from_mx(plhs[i]))
                L.mock_destroy(plhs[i])
        return outs[0] if nlhs == 1 else outs


class Ptr:
    def __init__(self, p):
        self.p = p


def desc(wn, sizes):
    d = len(sizes)
    names = [wn] * d if isinstance(wn, str) else list(wn)
    return {"wname": names, "sizes": np.asarray([[float(s) for s in sizes]])}


@pytest.fixture(scope="module", params=["interleaved", "split"])
def mex(request):
    m = Mex(request.param)
    yield m
    m.call("release", nlhs=0)


# ------------------------------------------------------------------------------------------------ no device needed
def test_gateway_taps_and_error_texts(mex):
    lo, hi = mex.call("taps", "db4", nlhs=2)
    olo, ohi = orc.wave_filters("db4")
    assert lo.shape == (1, 8) and np.array_equal(lo.ravel(), olo) and np.array_equal(hi.ravel(), ohi)
    with pytest.raises(MexError, match="Unknown Wavelet Name"):
        mex.call("taps", "db11", nlhs=2)
    with pytest.raises(MexError, match="unknown nd_dwt_mex command"):
        mex.call("frobnicate", nlhs=0)
    x = orc.synth((16, 12), np.complex128, 1)
    with pytest.raises(MexError, match="Four Inputs Required") as ei:          # mex/nd_dwt_mex.c:19-22
        mex.call(x, desc("db2", (16, 12)), 0.0)
    assert ei.value.ident == "MATLAB:FFT2mx:invalidNumInputs"
    with pytest.raises(MexError, match="level must be in 1..16"):
        mex.call(x, desc("db2", (16, 12)), 0.0, 0.0, 0.0)
    with pytest.raises(MexError, match="FIlter size and image size not consistant"):   # not a descriptor / handle
        mex.call(x, np.zeros((2, 2)), 0.0, 1.0, 0.0)
    with pytest.raises(MexError, match="FIlter size and image size not consistant"):   # wname / sizes disagree
        mex.call(x, {"wname": ["db2"], "sizes": np.asarray([[16.0, 12.0]])}, 0.0, 1.0, 0.0)
    with pytest.raises(MexError, match="Arrays must be double or single"):
        mex.call(np.zeros((16, 12), dtype=np.uint64), desc("db2", (16, 12)), 0.0, 1.0, 0.0)
    with pytest.raises(MexError, match="not a plan handle"):
        mex.call("shrink", np.asarray([[7.0]]), np.zeros((1, 4)), nlhs=0)
    assert mex.L.mock_live_mallocs() == 0          # every mxArrayToString / mxMalloc was released, also on the error paths


# ------------------------------------------------------------------------------------------------ on the B200
CASES = [((300,), "db3", 3, np.complex128, 0), ((48, 40), ["db1", "db4"], 2, np.complex64, 1), ((24, 20, 16), "db2", 2, np.float64, 0),
         ((16, 12, 10, 8), "db1", 2, np.complex64, 0), ((64, 48), "db4", 3, np.float32, 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("sizes,wn,level,dtype,l2", CASES)
def test_gateway_plan_handle_dec_rec(mex, sizes, wn, level, dtype, l2):
    """h = nd_dwt_mex('plan', f, is_single, is_complex, l2);  y = nd_dwt_mex(x, h, 0, level, l2);  x = nd_dwt_mex(y, h, 1, ...)"""
    x = orc.synth(sizes, dtype, 5)
    single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
    tol = 1e-5 if single else 1e-12
    h = mex.call("plan", desc(wn, sizes), float(single), float(np.iscomplexobj(x)), float(l2))
    assert h.dtype == np.uint64 and h.shape == (1, 1)
    y = mex.call(x, h, 0.0, float(level), float(l2))
    nb = orc.num_bands(len(sizes), level)
    assert y.shape == tuple(sizes) + (nb,) and y.dtype == x.dtype            # [sizes, nb], mex/nd_dwt_mex.c:79-88
    yo = orc.dec_direct(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), wn, level, bool(l2))
    assert orc.rel_l2(y, yo) <= tol
    xr = mex.call(y, h, 1.0, float(level), float(l2))
    assert xr.shape == (tuple(sizes) if len(sizes) > 1 else (sizes[0], 1))    # 1-D: a column vector, nd_dwt_mex.c:136-138
    assert orc.rel_l2(xr.reshape(sizes), x) <= tol
    # the descriptor struct in the f position (what the reference's objects store as f_dec) finds / makes a plan by key
    y2 = mex.call(x, desc(wn, sizes), 0.0, float(level), float(l2))
    assert np.array_equal(y2, y)
    with pytest.raises(MexError, match="FIlter size and image size not consistant"):    # nd_dwt_mex.c:36-51
        mex.call(x.ravel()[:-1].copy(), h, 0.0, float(level), float(l2))
    other = x.astype(np.complex64 if x.dtype == np.complex128 else np.float64 if x.dtype == np.float32 else np.complex128)
    with pytest.raises(MexError, match="plan and data differ in class or complexity"):
        mex.call(other, h, 0.0, float(level), float(l2))
    mex.call("release", h, nlhs=0)
    with pytest.raises(MexError):                                            # a released handle is no plan any more
        mex.call(x, h, 0.0, float(level), float(l2))
    assert mex.L.mock_live_mallocs() == 0


@pytest.mark.gpu
def test_gateway_shrink_dilations_release(mex):
    sizes, level = (40, 32, 12), 2
    x = orc.synth(sizes, np.complex64, 6)
    h = mex.call("plan", desc("db2", sizes), 1.0, 1.0, 0.0)
    tab = np.array([[0.0, 0.3, 0.1, 0.6, 0.2, 0.9, 0.4, 0.5], [0.0, 1.0, 0.7, 0.2, 0.8, 0.3, 0.6, 0.1]])   # table(j, b)
    mex.call("shrink", h, tab, nlhs=0)
    y = mex.call(x, h, 0.0, float(level), 0.0)
    ref = orc.shrink_soft(orc.dec_direct(x.astype(np.complex128), "db2", level), tab, 3)
    assert orc.rel_l2(y, ref) <= 1e-5
    with pytest.raises(MexError, match="threshold table"):
        mex.call("shrink", h, np.zeros((2, 5)), nlhs=0)
    mex.call("shrink", h, np.zeros((0, 0)), nlhs=0)                           # [] switches the threshold off
    y = mex.call(x, h, 0.0, float(level), 0.0)
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), "db2", level)) <= 1e-5
    mex.call("dilations", h, np.asarray([[1.0, 2.0]]), nlhs=0)                # a-trous option
    y = mex.call(x, h, 0.0, float(level), 0.0)
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), "db2", level, dilations=[1, 2])) <= 1e-5
    assert orc.rel_l2(mex.call(y, h, 1.0, float(level), 0.0), x) <= 1e-5
    mex.call("release", nlhs=0)                                              # everything
    with pytest.raises(MexError):
        mex.call(x, h, 0.0, float(level), 0.0)
    mex.L.mock_run_atexit()                                                  # what MATLAB calls when the MEX file is cleared


@pytest.mark.gpu
def test_gateway_gpuarray_branch_stays_on_device():
    """mexcuda flavour: a gpuArray goes in by device pointer and a gpuArray comes out (no gather / upload per call, the
    reference's 'gpu_off' traffic, Functions/nd_dwt_1D.m:139-141,192-194); host arrays still work in the same build."""
    import torch
    m = Mex("gpu")
    L = m.L
    for sizes, wn, level, dtype in [((40, 36, 20), "db4", 2, np.complex64), ((24, 20, 12, 16), "db4", 2, np.complex64),
                                    ((96, 80), ["db2", "db3"], 3, np.complex128), ((5000,), "db8", 4, np.float32)]:
        x = orc.synth(sizes, dtype, 7)
        single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
        tol = 1e-5 if single else 1e-12
        h = m.call("plan", desc(wn, sizes), float(single), float(np.iscomplexobj(x)), 0.0)
        xh = m.to_mx(x)
        xg = L.mock_gpu_array(xh)
        assert xg and L.mock_is_gpu(xg)
        yg = m.call(Ptr(xg), h, 0.0, float(level), 0.0, raw_out=True)
        assert L.mock_is_gpu(yg.p)                                           # the result is a gpuArray
        yh = L.mock_gather(yg.p)
        y = m.from_mx(yh)
        yo = orc.dec_direct(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), wn, level)
        assert y.shape == tuple(sizes) + (orc.num_bands(len(sizes), level),) and orc.rel_l2(y, yo) <= tol
        xrg = m.call(yg, h, 1.0, float(level), 0.0, raw_out=True)
        xrh = L.mock_gather(xrg.p)
        assert orc.rel_l2(m.from_mx(xrh).reshape(sizes), x) <= tol
        assert np.array_equal(m.from_mx(L.mock_gather(xg)), m.from_mx(xh))    # the input gpuArray is untouched
        with pytest.raises(MexError, match="FIlter size and image size not consistant"):
            m.call(Ptr(yg.p), h, 0.0, float(level), 0.0)                     # a stack where an array is expected
        assert orc.rel_l2(m.call(x, h, 0.0, float(level), 0.0), yo) <= tol   # host arrays through the same build
        for p in (xh, xg, yg.p, yh, xrg.p, xrh):
            L.mock_destroy(p)
        m.call("release", h, nlhs=0)
    assert L.mock_live_gpu_handles() == 0 and L.mock_live_mallocs() == 0     # no mxGPUArray handle left alive, also after errors
    if torch.cuda.device_count() >= 2:                                       # 'ngpus' plans take host arrays and use every GPU
        sizes, level = (32, 24, 16, 12), 2
        x = orc.synth(sizes, np.complex64, 8)
        h = m.call("plan", desc("db4", sizes), 1.0, 1.0, 0.0, 2.0)
        y = m.call(x, h, 0.0, float(level), 0.0)
        assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), "db4", level)) <= 1e-5
        assert orc.rel_l2(m.call(y, h, 1.0, float(level), 0.0), x) <= 1e-5
        xg = L.mock_gpu_array(m.to_mx(x))
        with pytest.raises(MexError, match="multi-GPU plans take host arrays"):
            m.call(Ptr(xg), h, 0.0, float(level), 0.0)
    m.call("release", nlhs=0)
