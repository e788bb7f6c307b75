"""Host-side mirror of the reference's operator interface for the dec/rec hot path.

Same names, argument meaning and error behaviour as the MATLAB objects of
arg-min-x/Non-Decimated_Wavelets (Functions/nd_dwt_{1,2,3,4}D.m, harr_nddwt_{2,4}D.m,
wave_filters.m) and the `nd_dwt_mex` entry (mex/nd_dwt_mex.c), so that the parity tests read
like the reference's own Test/*.m scripts:

    obj = nd_dwt_3D({'db1','db3','db1'}, [164,64,40], 'pres_l2_norm', 1, 'compute', 'gpu')
    y   = obj.dec(x, level);   x2 = obj.rec(y)

Everything computes on the B200 through libnddwt_b200.so (ctypes -> C ABI); there is no CPU
path.  The `compute` option keeps the reference's four values but they now only pick where the
arrays live:  'mat' / 'mex' / 'gpu_off' take and return host (numpy) arrays -- the library copies
in and out, like the reference's 'gpu_off' (nd_dwt_1D.m:139-141,192-194);  'gpu' expects and
returns device-resident torch tensors (the reference's gpuArray).  Passing a CUDA tensor always
selects the device-resident path.

Array convention: shape == sizes (MATLAB order).  numpy arrays are handled in column-major
(Fortran) order without the caller having to care; torch tensors are expected as the
MATLAB-shaped view `t.permute(reversed dims)` of a C-contiguous tensor (see `to_device`).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib
from ._lib import NDDWT_C128, NDDWT_C64, NDDWT_F32, NDDWT_F64, NddwtError, Plan

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


# ------------------------------------------------------------------------------------------
def wave_filters(wname):
    """[low_d, hi_d] = wave_filters(wname) -- Functions/wave_filters.m:1.  Unknown names raise
    'Unknown Wavelet Name' (`:158-159`)."""
    try:
        return _lib.wave_filters_c(wname)
    except NddwtError as e:
        raise ValueError(str(e)) from None


def _np_dtype_code(precision, is_complex):
    single = str(precision).lower() == "single"
    if is_complex:
        return (NDDWT_C64, np.complex64) if single else (NDDWT_C128, np.complex128)
    return (NDDWT_F32, np.float32) if single else (NDDWT_F64, np.float64)


def _lib_max_levels():
    return 16      # NDDWT_MAX_LEVELS


_TORCH_OF = {}
if torch is not None:
    _TORCH_OF = {NDDWT_F32: torch.float32, NDDWT_F64: torch.float64,
                 NDDWT_C64: torch.complex64, NDDWT_C128: torch.complex128}


def to_device(x, device="cuda:0"):
    """numpy (MATLAB-shaped) -> device tensor whose memory is column-major; returned as the
    MATLAB-shaped permuted view."""
    a = np.asfortranarray(x)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(device)   # C-contiguous reversed shape
    return t.permute(*reversed(range(t.dim())))


def to_host(t):
    """inverse of to_device."""
    d = t.dim()
    return np.ascontiguousarray(t.permute(*reversed(range(d))).contiguous().cpu().numpy()).T


def _colmajor_base(t):
    """Return the C-contiguous reversed-shape tensor underlying a MATLAB-shaped tensor (copying
    only when the memory is not already column-major)."""
    r = t.permute(*reversed(range(t.dim())))
    return r if r.is_contiguous() else r.contiguous()


class _NdDwtBase:
    """Common constructor / dec / rec for nd_dwt_1D..4D (reference: one copy per class)."""

    _ndims = 0
    _extra_options = ()

    # -- constructor: nd_dwt_2D.m:78-138 (1-D :79-133, 3-D :80-139, 4-D :79-134) --------------
    def __init__(self, wname, sizes, *varargin, **kwargs):
        d = self._ndims
        sizes = [int(s) for s in np.atleast_1d(sizes).tolist()]
        if len(sizes) != d:
            if d == 1:
                raise ValueError("1D array length must be a scalar")
            raise ValueError("The sizes vector must be length %d" % d)
        self.sizes = tuple(sizes)
        if len(varargin) % 2:
            raise ValueError("Optional inputs must come in pairs")
        if isinstance(wname, str):
            self.wname = [wname] * d
        else:
            if d == 1:
                raise ValueError("Wavelet Name Must be a string")
            wname = list(wname)
            if len(wname) != d:
                raise ValueError("You must specify %d filter names in a cell array of length %d, or a single "
                                 "string for the same filter to be used in all dimensions" % (d, d))
            self.wname = wname
        self.pres_l2_norm = 0
        self.precision = "double"
        self.compute = "mat"
        self.method = "fft"
        self.ngpus = 1
        self.devices = None
        self.dilations = None
        opts = list(zip(varargin[0::2], varargin[1::2])) + list(kwargs.items())
        for ind, (key, val) in enumerate(opts):
            k = str(key).lower()
            if k == "pres_l2_norm":
                self.pres_l2_norm = int(bool(val))
            elif k == "compute":
                self.compute = str(val)
            elif k == "precision":
                self.precision = str(val)
            elif k == "atrous":        # extension: true a-trous transform, dilation 2^(j-1) at level j (the reference
                if val:                #            re-applies undilated filters, nd_dwt_2D.m:183)
                    self.dilations = [1 << j for j in range(_lib_max_levels())]
            elif k == "ngpus":         # extension: slabs of the last dimension over several GPUs (host arrays)
                self.ngpus = int(val)
            elif k == "devices":       # extension: explicit device list for 'ngpus' (repeat an index to emulate ranks)
                self.devices = [int(v) for v in val]
                self.ngpus = len(self.devices)
            elif k in self._extra_options:
                setattr(self, k, val)
            else:
                warnings.warn("Unknown optional input #%d ingoring!" % (2 * ind + 1))
        if self.compute.lower() not in ("mat", "mex", "gpu", "gpu_off"):
            raise ValueError("compute must be one of 'mat', 'mex', 'gpu', 'gpu_off'")
        # get_filters: taps + dimension check (nd_dwt_2D.m:259-277); errors surface here like the reference
        self.f_size = {}
        for i, w in enumerate(self.wname):
            lo, _ = wave_filters(w)
            self.f_size["s%d" % (i + 1)] = len(lo)
            if len(lo) > self.sizes[i]:
                raise ValueError("Dimension %d of Data is shorter than the wavelet filter being used" % (i + 1))
        self._plans = {}
        self.kernel_mode = 0
        self.params = {}
        self.shrink = None

    # -- plan cache (the stored-filter object on the device) ---------------------------------
    def _plan(self, is_complex, device_index, batch=1):
        code, _ = _np_dtype_code(self.precision, is_complex)
        key = (code, device_index, int(batch))
        pl = self._plans.get(key)
        if pl is None:
            pl = Plan(self.sizes, self.wname, code, self.pres_l2_norm, device_index)
            if batch != 1:
                pl.set_batch(batch)
            if self.dilations is not None:
                pl.set_dilations(self.dilations)
            pl.set_kernel_mode(self.kernel_mode)
            for name, value in self.params.items():
                pl.set_param(name, value)
            if self.shrink is not None:
                pl.set_shrink(self.shrink)
            self._plans[key] = pl
        return pl

    def set_shrink(self, thr):
        """Fused coefficient-domain shrink (extension; the step between dec and rec of the iterative loops the
        reference is meant for, README.md:2): every later `dec` returns soft-thresholded DETAIL coefficients
        (complex: c * max(0, 1 - t/|c|); real: sign(c) * max(|c| - t, 0)); the approximation band is exempt.
        thr: None (off), a scalar (one threshold for every detail band), a sequence of J values (one per level,
        finest first), or a table [J][2^d] (column 0 ignored)."""
        if thr is None:
            self.shrink = None
        else:
            nd = 1 << self._ndims
            t = np.asarray(thr, dtype=np.float64)
            if t.ndim == 0:
                t = np.full((_lib_max_levels(), nd), float(t))
            elif t.ndim == 1:
                t = np.repeat(t[:, None], nd, axis=1)
            elif t.ndim != 2 or t.shape[1] != nd:
                raise ValueError("thresholds: scalar, [J] or [J][2^d]")
            if np.any(t < 0):
                raise ValueError("thresholds must be >= 0")
            self.shrink = np.ascontiguousarray(t)
        for pl in self._plans.values():
            pl.set_shrink(self.shrink)

    def _mplan(self, is_complex):
        """Multi-GPU plan of this object (one process drives `ngpus` devices; host arrays in and out)."""
        code, _ = _np_dtype_code(self.precision, is_complex)
        key = ("multi", code)
        pl = self._plans.get(key)
        if pl is None:
            devs = self.devices if self.devices is not None else list(range(self.ngpus))
            pl = _lib.MultiPlan(self.sizes, self.wname, code, self.pres_l2_norm, devices=devs)
            if self.dilations is not None:      # halo buffers are sized by the largest dilation: keep the useful levels
                pl.set_dilations([d for d in self.dilations if d <= self.sizes[-1]] or [1])
            pl.set_kernel_mode(self.kernel_mode)
            for name, value in self.params.items():
                pl.set_param(name, value)
            if self.shrink is not None:
                pl.set_shrink(self.shrink)
            self._plans[key] = pl
        return pl

    def set_param(self, name, value):
        """Named integer plan parameters (nddwt_plan_set_param), e.g. 'rows_min_ctas'."""
        self.params[str(name)] = int(value)
        for pl in self._plans.values():
            pl.set_param(name, value)

    def set_dilations(self, dil):
        """Opt-in a-trous mode (not reference behaviour): dilation per level, finest first."""
        self.dilations = [int(v) for v in dil]
        for pl in self._plans.values():
            pl.set_dilations(self.dilations)

    def set_kernel_mode(self, mode):
        self.kernel_mode = int(mode)
        for pl in self._plans.values():
            pl.set_kernel_mode(self.kernel_mode)

    def _num_bands(self, level):
        nd = 1 << self._ndims
        return nd + (nd - 1) * (level - 1)

    def _level_of(self, nb):
        nd = 1 << self._ndims
        lv = 1 + (nb - nd) / (nd - 1)
        if lv < 1 or int(lv) != lv:
            raise ValueError("FIlter size and image size not consistant")
        return int(lv)

    def _check_level(self, level):
        if int(level) != level or level < 1:
            raise ValueError("level must be a positive integer")
        return int(level)

    # -- dec: nd_dwt_2D.m:141-197 ---------------------------------------------------------
    def dec(self, x, level, out=None):
        """y = obj.dec(x, level).  `out` (optional, host path): a preallocated column-major array
        of shape [sizes, nb] to write into (e.g. pinned memory reused across an iterative loop)."""
        level = self._check_level(level)
        if torch is not None and isinstance(x, torch.Tensor) and x.is_cuda:
            return self._dec_device(x, level)
        if self.compute.lower() == "gpu" and not (torch is not None and isinstance(x, torch.Tensor)):
            x = to_device(np.asarray(x))
            return self._dec_device(x, level)
        return self._dec_host(np.asarray(x), level, out)

    def rec(self, y, out=None):
        if torch is not None and isinstance(y, torch.Tensor) and y.is_cuda:
            return self._rec_device(y)
        if self.compute.lower() == "gpu" and not (torch is not None and isinstance(y, torch.Tensor)):
            return self._rec_device(to_device(np.asarray(y)))
        return self._rec_host(np.asarray(y), out)

    def _check_x_shape(self, shape):
        """Returns the batch count: 1 for the reference's shapes, B for the batched extension
        (one extra trailing dimension: x is [sizes, B], coefficients [sizes, B, nb])."""
        shape = tuple(shape)
        if shape == self.sizes:
            return 1
        if len(shape) == self._ndims + 1 and shape[:-1] == self.sizes:
            return int(shape[-1])
        raise ValueError("FIlter size and image size not consistant")

    @staticmethod
    def _check_out(out, shape, npdt):
        if out.shape != tuple(shape) or out.dtype != npdt or not out.flags.f_contiguous:
            raise ValueError("out must be a column-major array of shape %s and dtype %s" % (tuple(shape), npdt))
        return out

    def _dec_host(self, x, level, out=None):
        if self._ndims == 1 and x.ndim == 2 and 1 in x.shape:
            x = x.reshape(-1)
        batch = self._check_x_shape(x.shape)
        is_c = np.iscomplexobj(x)
        code, npdt = _np_dtype_code(self.precision, is_c)
        xf = np.asfortranarray(x, dtype=npdt)
        shape = tuple(x.shape) + (self._num_bands(level),)
        y = np.empty(shape, dtype=npdt, order="F") if out is None else self._check_out(out, shape, np.dtype(npdt))
        if self.ngpus > 1 and batch == 1:
            self._mplan(is_c).dec_host(xf.ctypes.data, y.ctypes.data, level)
        else:
            self._plan(is_c, 0, batch).dec_host(xf.ctypes.data, y.ctypes.data, level)
        return y

    def _rec_host(self, y, out=None):
        if y.ndim not in (self._ndims + 1, self._ndims + 2):
            raise ValueError("FIlter size and image size not consistant")
        batch = self._check_x_shape(y.shape[:-1])
        level = self._level_of(y.shape[-1])
        is_c = np.iscomplexobj(y)
        code, npdt = _np_dtype_code(self.precision, is_c)
        yf = np.asfortranarray(y, dtype=npdt)
        xs = tuple(y.shape[:-1])
        x = np.empty(xs, dtype=npdt, order="F") if out is None else self._check_out(out, xs, np.dtype(npdt))
        if self.ngpus > 1 and batch == 1:
            self._mplan(is_c).rec_host(yf.ctypes.data, x.ctypes.data, level)
        else:
            self._plan(is_c, 0, batch).rec_host(yf.ctypes.data, x.ctypes.data, level)
        return x

    def _dec_device(self, x, level):
        batch = self._check_x_shape(x.shape)
        is_c = x.is_complex()
        code, _ = _np_dtype_code(self.precision, is_c)
        tdt = _TORCH_OF[code]
        base = _colmajor_base(x.to(tdt))
        dev = x.device.index or 0
        out = torch.empty((self._num_bands(level),) + tuple(reversed(tuple(x.shape))), dtype=tdt, device=x.device)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._plan(is_c, dev, batch).dec(base.data_ptr(), out.data_ptr(), level, stream)
        return out.permute(*reversed(range(out.dim())))

    def _rec_device(self, y):
        if y.dim() not in (self._ndims + 1, self._ndims + 2):
            raise ValueError("FIlter size and image size not consistant")
        batch = self._check_x_shape(y.shape[:-1])
        level = self._level_of(y.shape[-1])
        is_c = y.is_complex()
        code, _ = _np_dtype_code(self.precision, is_c)
        tdt = _TORCH_OF[code]
        base = _colmajor_base(y.to(tdt))
        dev = y.device.index or 0
        out = torch.empty(tuple(reversed(tuple(y.shape[:-1]))), dtype=tdt, device=y.device)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._plan(is_c, dev, batch).rec(base.data_ptr(), out.data_ptr(), level, stream)
        return out.permute(*reversed(range(out.dim())))

    def launches(self):
        return sum(pl.launches for pl in self._plans.values())

    def synthesis_kernels(self):
        """Synthesis tile kernel each plan of this object launched last (see Plan.last_synthesis_kernel)."""
        return sorted(set(pl.last_synthesis_kernel for pl in self._plans.values() if hasattr(pl, "last_synthesis_kernel")))


class nd_dwt_1D(_NdDwtBase):
    """Functions/nd_dwt_1D.m:63-321."""
    _ndims = 1


class nd_dwt_2D(_NdDwtBase):
    """Functions/nd_dwt_2D.m:64-339."""
    _ndims = 2


class nd_dwt_3D(_NdDwtBase):
    """Functions/nd_dwt_3D.m:66-395."""
    _ndims = 3


class nd_dwt_4D(_NdDwtBase):
    """Functions/nd_dwt_4D.m:65-469 ('method' is accepted like the reference, `:74,97,111-112`;
    both values run the same direct spatial kernels here)."""
    _ndims = 4
    _extra_options = ("method",)


class _HaarBase(_NdDwtBase):
    """harr_nddwt_2D.m / harr_nddwt_4D.m: Haar == db1 in every dimension, level 1."""

    _multi_level_ok = False

    def __init__(self, sizes, *varargin, **kwargs):
        super().__init__("db1", sizes, *varargin, **kwargs)
        # obj.scale of the reference (harr_nddwt_2D.m:122-126)
        self.scale = 0.5 if self.pres_l2_norm else 1.0 / np.sqrt(2.0)

    def dec(self, x, level=1):
        if level != 1 and not self._multi_level_ok:
            raise ValueError("Only single level decomposition supported for Harr")   # harr_nddwt_2D.m:136-138
        return super().dec(x, level)


class harr_nddwt_2D(_HaarBase):
    """Functions/harr_nddwt_2D.m:64-325 (level 1 only, `:136-138`)."""
    _ndims = 2


class harr_nddwt_4D(_HaarBase):
    """Functions/harr_nddwt_4D.m:65-884.  The reference has no level check but its level>1 path is
    not invertible; here level>1 is defined as the db1 multi-level transform."""
    _ndims = 4
    _multi_level_ok = True


_CLASS_OF_DIM = {1: nd_dwt_1D, 2: nd_dwt_2D, 3: nd_dwt_3D, 4: nd_dwt_4D}


class FilterSpec:
    """What travels in nd_dwt_mex's `f_dec` position.  The reference passes the stored
    Fourier-domain filters (2^d * numel complex values, nd_dwt_2D.m:160); the spatial kernels only
    need the wavelet names and sizes, so the new classes pass this small descriptor instead."""

    def __init__(self, wname, sizes, precision="double"):
        self.sizes = tuple(int(s) for s in np.atleast_1d(sizes).tolist())
        d = len(self.sizes)
        self.wname = [wname] * d if isinstance(wname, str) else list(wname)
        self.precision = precision
        self._objs = {}

    def obj(self, pres_l2):
        key = int(bool(pres_l2))
        if key not in self._objs:
            cls = _CLASS_OF_DIM[len(self.sizes)]
            wn = self.wname[0] if len(self.sizes) == 1 else self.wname
            self._objs[key] = cls(wn, self.sizes, "pres_l2_norm", key, "precision", self.precision, "compute", "mex")
        return self._objs[key]


def nd_dwt_mex(x, f_dec, direction, level, pres_l2_norm, *rest):
    """y = nd_dwt_mex(x, f_dec, dir, level, pres_l2_norm) -- the five-argument entry of
    mex/nd_dwt_mex.c:8-153.  dir == 0: forward ([sizes] -> [sizes, nb]); otherwise inverse.
    `x` is the SPATIAL array (the reference passes its FFT; the FFT was an implementation detail
    of the fast convolution) and `f_dec` a FilterSpec.  Errors mirror the gateway's
    (`:19-30,36-51,124-127`)."""
    if not isinstance(f_dec, FilterSpec):
        raise TypeError("Four Inputs Required")   # the gateway's only arity message (nd_dwt_mex.c:19-22)
    obj = f_dec.obj(pres_l2_norm)
    if direction == 0:
        return obj.dec(x, int(level))
    y = obj.rec(x)
    return y
