"""CPU oracle for the non-decimated wavelet dec/rec hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the algorithm of the reference
(arg-min-x/Non-Decimated_Wavelets) for the path BASELINE.json names.  It is the checker the
parity tests, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs use; nothing in the product package imports it, and the product has no CPU fallback.

Pinning status: the reference holds NO golden vectors or known-answer tests for this path
(its Test/*.m and mex/*_test.m scripts only print PR error / energy, with unseeded randn).
The oracle is therefore pinned by
  (1) `oracle/_ref` -- the reference's own `mex/nddwt.c`, compiled from where it lies under
      /root/reference and linked against a stand-in for the five FFTW symbols it calls
      (oracle/fftw_standin.c); `tests/test_oracle.py` feeds both the same FFT-domain
      inputs and compares (runs in the build container, where /root/reference exists);
  (2) the properties the reference's scripts print: perfect reconstruction, energy
      preservation with pres_l2_norm, mat-path == mex-path, Haar == db1;
  (3) an independent closed-form spatial restatement (`dec_direct` / `rec_direct`);
  (4) the tap table re-derived by spectral factorisation (oracle/gen_taps.py) and compared with
      the reference's literals.
The MATLAB classes themselves cannot be executed here (no MATLAB/Octave), so the 'mat' path is
restated, not run.

Array convention: `x.shape == sizes` (MATLAB order, dim 1 first), coefficients
`y.shape == sizes + (nb,)`, nb = 1 + J*(2^d - 1), deepest level first.  Memory order is
irrelevant here; numpy axis i is MATLAB dim i+1.
"""
from __future__ import annotations

import itertools
import numpy as np

try:  # scipy.fft is threaded (workers=) and keeps complex64; numpy.fft upcasts to complex128
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None

from .taps_table import DB_TAPS

_WORKERS = 1


def set_fft_workers(n: int) -> None:
    """Threads scipy.fft may use (the reference hard-codes 8 FFTW threads, mex/nddwt.c:103)."""
    global _WORKERS
    _WORKERS = max(1, int(n))


def _fftn(a, axes):
    if _sfft is not None:
        return _sfft.fftn(a, axes=axes, workers=_WORKERS)
    return np.fft.fftn(a, axes=axes)


def _ifftn(a, axes):
    if _sfft is not None:
        return _sfft.ifftn(a, axes=axes, workers=_WORKERS)
    return np.fft.ifftn(a, axes=axes)


# --------------------------------------------------------------------------------------
# wave_filters  (Functions/wave_filters.m:19-172)
# --------------------------------------------------------------------------------------
def wave_filters(wname: str):
    """[low_d, hi_d] = wave_filters(wname)  -- Functions/wave_filters.m:1.

    Table value h (`:21-156`); hi = flip(h) with 1-based even entries negated (`:164-166`),
    then both flipped (`:171-172`).  Unknown name -> error (`:158-159`).
    """
    key = str(wname).lower()
    if not (key.startswith("db") and key[2:].isdigit() and int(key[2:]) in DB_TAPS):
        raise ValueError("Unknown Wavelet Name")
    h = np.asarray(DB_TAPS[int(key[2:])], dtype=np.float64)
    hi = h[::-1].copy()
    hi[1::2] = -hi[1::2]          # hi_d(2:2:end) = -hi_d(2:2:end)
    low_d = h[::-1].copy()
    hi_d = hi[::-1].copy()
    return low_d, hi_d


def _as_wnames(wname, d):
    if isinstance(wname, str):
        return [wname] * d
    wname = list(wname)
    if len(wname) != d:
        raise ValueError("need one wavelet name per dimension")
    return wname


def num_bands(d: int, level: int) -> int:
    return (1 << d) + ((1 << d) - 1) * (level - 1)   # mex/nd_dwt_mex.c:83


def infer_level(d: int, nb: int) -> int:
    """Level from the band count: nd_dwt_1D.m:213, nd_dwt_2D.m:215, nd_dwt_4D.m:213."""
    lv = 1 + (nb - (1 << d)) // ((1 << d) - 1)
    if num_bands(d, lv) != nb:
        raise ValueError("band count does not match any level")
    return lv


# --------------------------------------------------------------------------------------
# 'mat' path: stored Fourier-domain filters, fast convolution
# --------------------------------------------------------------------------------------
def get_filters(wname, sizes, pres_l2_norm=False, mex_scale=False, dtype=np.complex128):
    """f_dec[..., b] of get_filters: nd_dwt_1D.m:257-290, nd_dwt_2D.m:259-309,
    nd_dwt_3D.m:263-342, nd_dwt_4D.m:255-391.

    Band b = sum_i b_i 2^i (dim 1 is the LSB; 0 = low-pass).  scale = 2^(-d/2) with
    pres_l2_norm, scale2 = 1/prod(sizes) only for the MEX path (its inverse DFT is
    unnormalised).  Phase ramp exp(+j 2 pi (L_i/2) k_i / N_i) per dim.
    """
    sizes = tuple(int(s) for s in sizes)
    d = len(sizes)
    wn = _as_wnames(wname, d)
    filt = [wave_filters(w) for w in wn]
    for i, (lo, _) in enumerate(filt):
        if len(lo) > sizes[i]:
            raise ValueError("Dimension %d of Data is shorter than the wavelet filter being used" % (i + 1))
    scale = 2.0 ** (-d / 2.0) if pres_l2_norm else 1.0
    scale2 = 1.0 / float(np.prod(sizes)) if mex_scale else 1.0
    # per-dim 1-D spectra with the half-length phase ramp
    spec = []
    for i in range(d):
        lo, hi = filt[i]
        n = sizes[i]
        phase = np.exp(1j * 2 * np.pi * (len(lo) / 2.0) * (np.arange(n) / n))
        spec.append((phase * np.fft.fft(lo, n), phase * np.fft.fft(hi, n)))
    f_dec = np.empty(sizes + (1 << d,), dtype=np.complex128)
    for b in range(1 << d):
        acc = np.ones((), dtype=np.complex128)
        for i in range(d):
            shape = [1] * d
            shape[i] = sizes[i]
            acc = acc * spec[i][(b >> i) & 1].reshape(shape)
        f_dec[..., b] = scale2 * scale * acc
    return f_dec.astype(dtype)


def level_1_dec(x_f, f_dec):
    """y(..,b) = ifftn(x_f .* f_dec(..,b))  -- nd_dwt_2D.m:312-331 (1-D :293-310, 3-D :345-368, 4-D :394-424)."""
    d = x_f.ndim
    axes = tuple(range(d))
    return _ifftn(x_f[..., None] * f_dec, axes)


def level_1_rec(c_f, f_dec):
    """ifftn(sum_b c_f(..,b) .* conj(f_dec(..,b)))  -- nd_dwt_2D.m:334-337 (1-D :313-318, 3-D :371-393, 4-D :447-467)."""
    d = c_f.ndim - 1
    return _ifftn(np.sum(c_f * np.conj(f_dec), axis=-1), tuple(range(d)))


def dec(x, wname, level, pres_l2_norm=False, precision="double"):
    """obj.dec(x, level), compute='mat'  -- nd_dwt_2D.m:141-197 and siblings."""
    x = np.asarray(x)
    d = x.ndim
    cdt = np.complex64 if precision == "single" else np.complex128
    f_dec = get_filters(wname, x.shape, pres_l2_norm, dtype=cdt)
    x_real = not np.iscomplexobj(x)
    axes = tuple(range(d))
    y = level_1_dec(_fftn(x.astype(cdt), axes), f_dec)
    for _ in range(1, level):
        # y = cat(d+1, level_1_dec(fftn(y(..,1))), y(..,2:end))
        y = np.concatenate([level_1_dec(_fftn(y[..., 0], axes), f_dec), y[..., 1:]], axis=-1)
    if x_real:
        y = y.real
    return y


def rec(y, wname, pres_l2_norm=False, precision="double"):
    """obj.rec(y), compute='mat'  -- nd_dwt_2D.m:200-252 and siblings."""
    y = np.asarray(y)
    d = y.ndim - 1
    nd = 1 << d
    level = infer_level(d, y.shape[-1])
    cdt = np.complex64 if precision == "single" else np.complex128
    f_dec = get_filters(wname, y.shape[:-1], pres_l2_norm, dtype=cdt)
    y_real = not np.iscomplexobj(y)
    axes = tuple(range(d))
    c_f = _fftn(y.astype(cdt), axes)
    out = None
    for ind in range(1, level + 1):
        if ind == 1:
            out = level_1_rec(c_f[..., :nd], f_dec)
        else:
            lo = nd + (ind - 2) * (nd - 1)
            stack = np.concatenate([_fftn(out, axes)[..., None], c_f[..., lo:lo + nd - 1]], axis=-1)
            out = level_1_rec(stack, f_dec)
        if not pres_l2_norm:
            out = out / nd
    if y_real:
        out = out.real
    return out


# --------------------------------------------------------------------------------------
# MEX flow restated (mex/nddwt.c:98-292, mex/nd_dwt_mex.c:79-83): Fourier-domain in, slot
# arithmetic, swap-trick unnormalised inverse DFT (scale2 folded into the filters).
# --------------------------------------------------------------------------------------
def mex_forward(x_f, f_dec_mex, level):
    """nd_dwt_mex(x_f, f_dec, 0, level, .) -- inputs: fftn(x) and the scale2-scaled filters."""
    d = x_f.ndim
    nd = 1 << d
    numel = x_f.size
    axes = tuple(range(d))
    out = np.zeros(x_f.shape + (num_bands(d, level),), dtype=np.complex128)

    def dec_1level(img_f):
        # pointByPoint + unnormalised inverse DFT (nddwt.c:123-128)
        return _ifftn(img_f[..., None] * f_dec_mex, axes) * numel

    start = (nd - 1) * (level - 1)                    # nddwt.c:210
    out[..., start:start + nd] = dec_1level(x_f)
    approx_f = _fftn(out[..., start], axes)          # :216-222
    for level_ind in range(level - 1, 0, -1):          # :225-234
        start = (nd - 1) * (level_ind - 1)
        out[..., start:start + nd] = dec_1level(approx_f)
        approx_f = _fftn(out[..., start], axes)
    return out


def mex_inverse(c_f, f_dec_mex, level, pres_l2_norm):
    """nd_dwt_mex(c_f, f_dec, 1, level, pres_l2) -- c_f = per-band fftn of the coefficients."""
    d = c_f.ndim - 1
    nd = 1 << d
    numel = int(np.prod(c_f.shape[:-1]))
    axes = tuple(range(d))
    c_f = c_f.copy()                                    # the reference mutates its input (:163)

    def rec_1level(bands_f):
        t = _ifftn(bands_f * np.conj(f_dec_mex), axes) * numel   # :163-164
        o = t.sum(axis=-1)                                        # :168-173
        if not pres_l2_norm:
            o = o * (1.0 / nd)                                    # :176-182
        return o

    out = rec_1level(c_f[..., 0:nd])                   # :256
    if level > 1:
        c_f[..., nd - 1] = _fftn(out, axes)            # :259-268
        for level_ind in range(2, level + 1):          # :271-290
            start = (nd - 1) * (level_ind - 1)
            out = rec_1level(c_f[..., start:start + nd])
            if level_ind != level:
                c_f[..., (nd - 1) * level_ind] = _fftn(out, axes)
    return out


def dec_mex(x, wname, level, pres_l2_norm=False):
    """obj.dec(x, level) with compute='mex' (double complex only): nd_dwt_2D.m:156-160."""
    x = np.asarray(x)
    d = x.ndim
    f = get_filters(wname, x.shape, pres_l2_norm, mex_scale=True)
    y = mex_forward(_fftn(x.astype(np.complex128), tuple(range(d))), f, level)
    return y.real if not np.iscomplexobj(x) else y


def rec_mex(y, wname, pres_l2_norm=False):
    y = np.asarray(y)
    d = y.ndim - 1
    level = infer_level(d, y.shape[-1])
    f = get_filters(wname, y.shape[:-1], pres_l2_norm, mex_scale=True)
    out = mex_inverse(_fftn(y.astype(np.complex128), tuple(range(d))), f, level, pres_l2_norm)
    return out.real if not np.iscomplexobj(y) else out


# --------------------------------------------------------------------------------------
# Closed-form spatial restatement -- independent of the FFT path
# --------------------------------------------------------------------------------------
def _filt_axis(a, g, axis, shift_sign, dil=1):
    """dec: out[n] = sum_k g[k] a[n - (k - L/2) dil];  rec (adjoint): out[n] = sum_k g[k] a[n + (k - L/2) dil]."""
    L = len(g)
    out = np.zeros_like(a)
    for k in range(L):
        off = (k - L // 2) * dil
        # np.roll(a, s)[n] = a[n - s]
        out += g[k] * np.roll(a, off if shift_sign > 0 else -off, axis=axis)
    return out


def dec_level_direct(a, filt, scale, dil=1):
    d = a.ndim
    bands = [a]
    # after processing axis i, list index bit i selects lo/hi of that axis
    for i in range(d):
        lo, hi = filt[i]
        nxt = [None] * (len(bands) * 2)
        for idx, arr in enumerate(bands):
            nxt[idx] = _filt_axis(arr, lo, i, +1, dil)
            nxt[idx + len(bands)] = _filt_axis(arr, hi, i, +1, dil)
        bands = nxt
    return [scale * b for b in bands]


def rec_level_direct(bands, filt, scale_norm, dil=1):
    d = bands[0].ndim
    cur = list(bands)
    for i in reversed(range(d)):
        lo, hi = filt[i]
        half = len(cur) // 2
        cur = [_filt_axis(cur[j], lo, i, -1, dil) + _filt_axis(cur[j + half], hi, i, -1, dil)
               for j in range(half)]
    return scale_norm * cur[0]


def dec_direct(x, wname, level, pres_l2_norm=False, dilations=None, dtype=None):
    """Spatial-domain multi-level analysis.  dilations=None -> 1 at every level (the reference's
    behaviour: the same f_dec is re-applied at each level, nd_dwt_2D.m:183, nddwt.c:214-228)."""
    x = np.asarray(x)
    if dtype is not None:
        x = x.astype(dtype)
    d = x.ndim
    nd = 1 << d
    filt = [wave_filters(w) for w in _as_wnames(wname, d)]
    for i, (lo, _) in enumerate(filt):
        if len(lo) > x.shape[i]:
            raise ValueError("Dimension %d of Data is shorter than the wavelet filter being used" % (i + 1))
    if x.dtype in (np.float32, np.complex64):
        filt = [(lo.astype(np.float32), hi.astype(np.float32)) for lo, hi in filt]
    scale = x.real.dtype.type(2.0 ** (-d / 2.0) if pres_l2_norm else 1.0)
    dil = [1] * level if dilations is None else list(dilations)
    out = np.empty(x.shape + (num_bands(d, level),), dtype=x.dtype)
    a = x
    for j in range(1, level + 1):
        bands = dec_level_direct(a, filt, scale, dil[j - 1])
        start = (nd - 1) * (level - j)
        for b in range(1, nd):
            out[..., start + b] = bands[b]
        a = bands[0]
    out[..., 0] = a
    return out


def rec_direct(y, wname, pres_l2_norm=False, dilations=None):
    y = np.asarray(y)
    d = y.ndim - 1
    nd = 1 << d
    level = infer_level(d, y.shape[-1])
    filt = [wave_filters(w) for w in _as_wnames(wname, d)]
    if y.dtype in (np.float32, np.complex64):
        filt = [(lo.astype(np.float32), hi.astype(np.float32)) for lo, hi in filt]
    scale = 2.0 ** (-d / 2.0) if pres_l2_norm else 1.0
    norm = 1.0 if pres_l2_norm else 1.0 / nd
    sn = y.real.dtype.type(scale * norm)
    dil = [1] * level if dilations is None else list(dilations)
    a = y[..., 0]
    for j in range(level, 0, -1):
        start = (nd - 1) * (level - j)
        bands = [a] + [y[..., start + b] for b in range(1, nd)]
        a = rec_level_direct(bands, filt, sn, dil[j - 1])
    return a


# --------------------------------------------------------------------------------------
# Haar classes, level 1 (harr_nddwt_2D.m:250-323, harr_nddwt_4D.m:248-882): slice add/sub
# --------------------------------------------------------------------------------------
def haar_scale(pres_l2_norm):
    return 0.5 if pres_l2_norm else 1.0 / np.sqrt(2.0)   # harr_nddwt_2D.m:122-126


def haar_level_1_dec(x, pres_l2_norm=False):
    """ap[n] = s (x[n] + x[n+1]), det[n] = s (x[n] - x[n+1]) periodic, dim by dim
    (harr_nddwt_2D.m:266-286; the 4-D class unrolls the same 16 pipelines, harr_nddwt_4D.m:268-551)."""
    x = np.asarray(x)
    s = x.real.dtype.type(haar_scale(pres_l2_norm))
    d = x.ndim
    bands = [x]
    for i in range(d):
        nxt = [None] * (2 * len(bands))
        for idx, arr in enumerate(bands):
            nb_ = np.roll(arr, -1, axis=i)               # x[n+1] with the wrap row (`:267,:270`)
            nxt[idx] = s * (arr + nb_)
            nxt[idx + len(bands)] = s * (arr - nb_)
        bands = nxt
    return np.stack(bands, axis=-1)


def haar_level_1_rec(y, pres_l2_norm=False):
    """s (c[n-1] + c[n]) for approximations, s (c[n] - c[n-1]) for details (harr_nddwt_2D.m:297-322),
    summed over bands; the caller divides by 2^d when not pres_l2_norm (`:221-223`)."""
    y = np.asarray(y)
    s = y.real.dtype.type(haar_scale(pres_l2_norm))
    d = y.ndim - 1
    cur = [y[..., b] for b in range(1 << d)]
    for i in reversed(range(d)):
        half = len(cur) // 2
        nxt = []
        for j in range(half):
            a, dt = cur[j], cur[j + half]
            nxt.append(s * (np.roll(a, 1, axis=i) + a) + s * (dt - np.roll(dt, 1, axis=i)))
        cur = nxt
    out = cur[0]
    if not pres_l2_norm:
        out = out / (1 << d)
    return out


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def synth(shape, dtype, seed):
    """Seeded synthetic input: N(0,1) (+ i N(0,1) for complex), cast to dtype."""
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if dt.kind == "c":
        r = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    else:
        r = rng.standard_normal(shape)
    return r.astype(dt)


def shrink_soft(y, table, ndims):
    """Soft threshold of the DETAIL bands of a coefficient stack `y` ([sizes, nb], deepest level first): the
    operation `nddwt_plan_set_shrink` fuses into the analysis kernels (extension; the
    reference leaves it to the caller's iterative loop, README.md:2).  table[j-1][b]: threshold of band b of
    level j (j = 1 finest), column 0 ignored.  Complex: c * max(0, 1 - t/|c|); real: sign(c) * max(|c| - t, 0)."""
    nd = 1 << ndims
    nb = y.shape[-1]
    level = 1 + (nb - nd) // (nd - 1)
    out = np.array(y, copy=True)
    table = np.asarray(table, dtype=np.float64)
    for j in range(1, level + 1):
        start = (nd - 1) * (level - j)           # slot arithmetic of mex/nddwt.c:209-210,226
        for b in range(1, nd):
            t = table[j - 1][b] if j - 1 < table.shape[0] else 0.0
            c = out[..., start + b]
            mag = np.abs(c)
            with np.errstate(divide="ignore", invalid="ignore"):
                sc = np.where(mag > t, 1.0 - t / np.where(mag > 0, mag, 1.0), 0.0)
            out[..., start + b] = c * sc
    return out
