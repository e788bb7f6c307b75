"""CPU tests of the oracle (test infrastructure) -- pins it against the reference's literals,
the reference's own compiled native core (oracle/_ref), the committed golden fixtures and the
properties the reference's Test/*.m scripts print."""
import glob
import os
import re

import numpy as np
import pytest

from oracle import nddwt_oracle as orc
from oracle import ref_mex
from conftest import REFERENCE_ROOT

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    wn = [str(s) for s in z["wname"]]
    return z["x"], z["y"], (wn[0] if len(wn) == 1 else wn), int(z["level"]), bool(z["pres_l2"])


def test_taps_match_reference_literals(have_reference):
    """P5: the re-derived Daubechies taps equal wave_filters.m:21-156 to the last double bit."""
    if not have_reference:
        pytest.skip("reference tree not present (GPU box)")
    src = open(os.path.join(REFERENCE_ROOT, "Functions", "wave_filters.m")).read()
    blocks = re.findall(r"case \{'db(\d+)'\}\s*low_d = \[(.*?)\];", src, re.S)
    assert len(blocks) == 10
    for p, body in blocks:
        p = int(p)
        if p == 1:
            continue
        vals = np.array([float(v) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+", body)])
        assert len(vals) == 2 * p
        assert np.max(np.abs(vals - np.array(orc.DB_TAPS[p]))) == 0.0


def test_wave_filters_known_answers():
    lo, hi = orc.wave_filters("db2")   # = MATLAB wfilters('db2') Lo_D / Hi_D
    r3, den = np.sqrt(3.0), 4 * np.sqrt(2.0)      # closed form of the D4 scaling filter
    h = np.array([1 + r3, 3 + r3, 3 - r3, 1 - r3]) / den
    np.testing.assert_allclose(lo, h[::-1], rtol=1e-14)
    np.testing.assert_allclose(hi, [-h[0], h[1], -h[2], h[3]], rtol=1e-14)
    for p in range(1, 11):
        lo, hi = orc.wave_filters("db%d" % p)
        assert len(lo) == 2 * p
        assert abs(lo.sum() - np.sqrt(2)) < 1e-14 and abs(hi.sum()) < 1e-13
        for m in range(p):   # orthonormality of even shifts
            assert abs(np.dot(lo[2 * m:], lo[:len(lo) - 2 * m]) - (m == 0)) < 1e-14
            assert abs(np.dot(lo[2 * m:], hi[:len(lo) - 2 * m])) < 1e-14
    with pytest.raises(ValueError, match="Unknown Wavelet Name"):
        orc.wave_filters("sym4")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    x, y, wn, level, l2 = _load(path)
    prec = "single" if x.dtype in (np.float32, np.complex64) else "double"
    tol = 1e-5 if prec == "single" else 1e-12
    assert orc.rel_l2(orc.dec(x, wn, level, l2, precision=prec), y) < tol
    assert orc.rel_l2(orc.dec_direct(x, wn, level, l2), y) < tol
    assert orc.rel_l2(orc.rec(y, wn, l2, precision=prec), x) < tol
    assert orc.rel_l2(orc.rec_direct(y, wn, l2), x) < tol


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_reference_core_reproduces_golden(path):
    """The reference's own mex/nddwt.c (compiled into oracle/_ref) reproduces the fixtures."""
    if not ref_mex.available():
        pytest.skip("oracle/_ref not built")
    x, y, wn, level, l2 = _load(path)
    tol = 1e-5 if x.dtype in (np.float32, np.complex64) else 1e-12
    assert orc.rel_l2(ref_mex.dec(x, wn, level, l2), y) < tol
    assert orc.rel_l2(ref_mex.rec(y, wn, l2), x) < tol


CASES = [
    ((54321,), "db1", 4, False),                                   # Test/nddwt1D_test.m:5-8
    ((264, 264), ["db1", "db3"], 1, True),                          # Test/nddwt2D_test.m:5-8
    ((41, 16, 10), ["db1", "db3", "db1"], 2, True),                 # Test/nddwt3D_test.m:5-7 (reduced)
    ((16, 16, 5, 6), ["db1", "db3", "db1", "db1"], 2, False),       # Test/nddwt4D_test.m:5-7 (reduced)
    ((16, 16, 10), ["db1", "db3", "db9"][:2] + ["db5"], 2, False),  # example_nd_dwt_3D-style mix (reduced)
]


@pytest.mark.parametrize("sizes,wn,level,l2", CASES)
def test_properties_P1_P2_P3(sizes, wn, level, l2):
    x = orc.synth(sizes, np.complex128, 7)
    y = orc.dec(x, wn, level, l2)
    assert y.shape == tuple(sizes) + (orc.num_bands(len(sizes), level),)
    assert orc.rel_l2(orc.rec(y, wn, l2), x) < 1e-12                      # P1 perfect reconstruction
    if l2:
        assert abs(np.linalg.norm(y) / np.linalg.norm(x) - 1) < 1e-12     # P2 energy
    assert orc.rel_l2(orc.dec_mex(x, wn, level, l2), y) < 1e-12           # P3 mat == mex flow
    assert orc.rel_l2(orc.rec_mex(y, wn, l2), x) < 1e-12
    assert orc.rel_l2(orc.dec_direct(x, wn, level, l2), y) < 1e-12        # closed form
    xr = orc.synth(sizes, np.float64, 8)
    yr = orc.dec(xr, wn, level, l2)
    assert not np.iscomplexobj(yr)                                        # real in -> real out


@pytest.mark.parametrize("sizes", [(10, 14), (6, 5, 4, 6)])
@pytest.mark.parametrize("l2", [False, True])
def test_haar_equals_db1_P4(sizes, l2):
    x = orc.synth(sizes, np.complex128, 3)
    yh = orc.haar_level_1_dec(x, l2)
    assert orc.rel_l2(yh, orc.dec(x, "db1", 1, l2)) < 1e-13
    assert orc.rel_l2(orc.haar_level_1_rec(yh, l2), x) < 1e-13


def test_impulse_response_and_adjoint():
    n, wn = 32, "db3"
    lo, hi = orc.wave_filters(wn)
    L = len(lo)
    x = np.zeros(n)
    x[0] = 1.0
    y = orc.dec(x, wn, 1)
    for k in range(L):      # dec(delta)[n] = g[n + L/2] at n in [-L/2, L/2-1] mod N
        assert abs(y[(k - L // 2) % n, 0] - lo[k]) < 1e-14
        assert abs(y[(k - L // 2) % n, 1] - hi[k]) < 1e-14
    a = orc.synth((12, 9), np.complex128, 1)
    c = orc.synth((12, 9, 4), np.complex128, 2)
    lhs = np.vdot(c, orc.dec(a, ["db2", "db3"], 1, True))
    rhs = np.vdot(orc.rec(c, ["db2", "db3"], True), a)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)      # <W a, c> = <a, W^T c>


def test_filter_longer_than_dim_errors():
    with pytest.raises(ValueError, match="shorter than the wavelet"):
        orc.dec(np.zeros((4, 32)), "db4", 1)


def test_single_precision_path_within_tolerance():
    x = orc.synth((24, 20, 12), np.complex64, 5)
    y = orc.dec(x, "db4", 3, precision="single")
    assert y.dtype == np.complex64
    y64 = orc.dec(x.astype(np.complex128), "db4", 3)
    assert orc.rel_l2(y, y64) < 1e-5
    assert orc.rel_l2(orc.rec(y, "db4", precision="single"), x) < 1e-5
