"""CPU tests: the C-ABI library loads, exports every symbol include/nddwt_b200.h declares, and
the host-side logic (taps, band arithmetic, argument checks, API option parsing) behaves like the
reference.  No compute calls (there is no GPU here and the library has no CPU path)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

import nddwt_b200 as nd
from oracle import nddwt_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nddwt_b200.h")).read()
    declared = set(re.findall(r"NDDWT_API[^;(]*?\b(nddwt_\w+)\s*\(", hdr))
    assert declared == set(nd._lib.SYMBOLS), declared ^ set(nd._lib.SYMBOLS)
    L = ctypes.CDLL(nd.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_wave_filters_matches_oracle():
    for p in range(1, 11):
        lo, hi = nd.wave_filters("db%d" % p)
        olo, ohi = orc.wave_filters("db%d" % p)
        assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
    lo, _ = nd.wave_filters("DB4")   # switch lower(wname), wave_filters.m:19
    assert len(lo) == 8
    with pytest.raises(ValueError, match="Unknown Wavelet Name"):
        nd.wave_filters("haar")


def test_band_arithmetic():
    L = nd.lib()
    for d in range(1, 5):
        for lv in range(1, 7):
            nb = L.nddwt_num_bands(d, lv)
            assert nb == orc.num_bands(d, lv)
            assert L.nddwt_infer_level(d, nb) == lv
        assert L.nddwt_infer_level(d, (1 << d) + 1) == (2 if d == 1 else 0)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    obj = nd.nd_dwt_2D("db2", [16, 16])
    with pytest.raises(nd.NddwtError, match="no usable CUDA device"):
        obj.dec(np.zeros((16, 16)), 1)


def test_constructor_option_parsing_and_errors():
    o = nd.nd_dwt_3D(["db1", "db3", "db1"], [164, 64, 40], "pres_l2_norm", 1, "compute", "gpu_off", "precision", "single")
    assert (o.pres_l2_norm, o.compute, o.precision) == (1, "gpu_off", "single")
    assert o.f_size == {"s1": 2, "s2": 6, "s3": 2}
    o = nd.nd_dwt_2D("db4", [256, 256])
    assert (o.pres_l2_norm, o.compute, o.precision, o.wname) == (0, "mat", "double", ["db4", "db4"])
    with pytest.raises(ValueError, match="sizes vector must be length 2"):
        nd.nd_dwt_2D("db1", [8, 8, 8])
    with pytest.raises(ValueError, match="scalar"):
        nd.nd_dwt_1D("db1", [8, 8])
    with pytest.raises(ValueError, match="Must be a string"):
        nd.nd_dwt_1D(["db1"], 64)
    with pytest.raises(ValueError, match="filter names"):
        nd.nd_dwt_3D(["db1", "db2"], [8, 8, 8])
    with pytest.raises(ValueError, match="come in pairs"):
        nd.nd_dwt_2D("db1", [8, 8], "pres_l2_norm")
    with pytest.raises(ValueError, match="Dimension 2 of Data is shorter"):
        nd.nd_dwt_2D(["db1", "db4"], [8, 6])
    with pytest.raises(ValueError, match="Unknown Wavelet Name"):
        nd.nd_dwt_2D("coif1", [8, 8])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        nd.nd_dwt_1D("db1", 64, "perserve_l2_norm", 1)   # the misspelt key of example_nd_dwt_1D.m:14
        assert any("Unknown optional input" in str(x.message) for x in w)
    o4 = nd.nd_dwt_4D("db1", [8, 8, 8, 8], "method", "conv")
    assert o4.method == "conv"


def test_haar_level_check():
    h = nd.harr_nddwt_2D([16, 16], "pres_l2_norm", 1)
    assert h.scale == 0.5
    with pytest.raises(ValueError, match="Only single level decomposition supported for Harr"):
        h.dec(np.zeros((16, 16)), 2)
    assert abs(nd.harr_nddwt_4D([4, 4, 4, 4]).scale - 1 / np.sqrt(2)) < 1e-16


def test_rec_rejects_bad_band_count():
    o = nd.nd_dwt_2D("db1", [8, 8])
    with pytest.raises(ValueError, match="not consistant"):
        o.rec(np.zeros((8, 8, 5)))
    with pytest.raises(ValueError, match="not consistant"):
        o.dec(np.zeros((8, 9)), 1)


def test_shrink_threshold_tables_host_side():
    """set_shrink accepts a scalar, one value per level, or a [J][2^d] table; negative values are refused.  (The
    table reaches the device plans at the next dec; no GPU is needed to build it.)"""
    o = nd.nd_dwt_3D("db2", [16, 16, 16], "precision", "single")
    o.set_shrink(0.5)
    assert o.shrink.shape == (16, 8) and np.all(o.shrink == 0.5)
    o.set_shrink([0.1, 0.2, 0.3])
    assert o.shrink.shape == (3, 8) and np.allclose(o.shrink[:, 5], [0.1, 0.2, 0.3])
    t = np.arange(16, dtype=float).reshape(2, 8)
    o.set_shrink(t)
    assert np.array_equal(o.shrink, t)
    o.set_shrink(None)
    assert o.shrink is None
    with pytest.raises(ValueError):
        o.set_shrink(-0.1)
    with pytest.raises(ValueError):
        o.set_shrink(np.zeros((2, 4)))      # a 3-D transform has 8 bands per level


def test_multi_gpu_options_and_export_blob_size():
    L = nd.lib()
    assert L.nddwt_mplan_export_size() >= 64 + 16          # a CUDA IPC handle + geometry
    o = nd.nd_dwt_4D("db4", [32, 24, 16, 12], "ngpus", 4)
    assert o.ngpus == 4 and o.devices is None
    o = nd.nd_dwt_4D("db4", [32, 24, 16, 12], "devices", [0, 0, 0])
    assert o.ngpus == 3 and o.devices == [0, 0, 0]
    # soft threshold of the oracle: identity at t = 0, approximation band exempt
    x = orc.synth((12, 10), np.complex128, 2)
    y = orc.dec_direct(x, "db2", 2)
    assert np.array_equal(orc.shrink_soft(y, np.zeros((2, 4)), 2), y)
    ys = orc.shrink_soft(y, np.full((2, 4), 1e9), 2)
    assert np.array_equal(ys[..., 0], y[..., 0]) and not ys[..., 1:].any()


def test_header_is_plain_c_and_every_entry_point_links(tmp_path):
    """The drop-in boundary is a C ABI: include/nddwt_b200.h compiles as C99 (no C++-isms, no torch types) and a C
    program that takes the address of every declared entry point links against libnddwt_b200.so; the entry points that
    need no device (wave_filters, num_bands, the error channel) answer from plain C exactly as through ctypes."""
    import subprocess
    hdr = open(os.path.join(ROOT, "include", "nddwt_b200.h")).read()
    names = sorted(set(re.findall(r"NDDWT_API[^;(]*?\b(nddwt_\w+)\s*\(", hdr)))
    src = tmp_path / "abi_check.c"
    src.write_text('#include "nddwt_b200.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t fn[] = {\n'
                   + "".join("    (fn_t)%s,\n" % n for n in names)
                   + '  };\n  double lo[20], hi[20]; int len = 0, rc, k;\n'
                     '  printf("%d entry points, %s\\n", (int)(sizeof fn / sizeof fn[0]), nddwt_version());\n'
                     '  rc = nddwt_wave_filters("db4", lo, hi, &len);\n  printf("rc %d len %d\\n", rc, len);\n'
                     '  for (k = 0; k < len; ++k) printf("%.17g %.17g\\n", lo[k], hi[k]);\n'
                     '  printf("nb %d\\n", (int)nddwt_num_bands(4, 3));\n'
                     '  rc = nddwt_wave_filters("db11", lo, hi, &len);\n  printf("rc %d msg %s\\n", rc, nddwt_last_error());\n'
                     '  return 0;\n}\n')
    libdir = os.path.dirname(nd.LIB_PATH)
    exe = tmp_path / "abi_check"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lnddwt_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].startswith("%d entry points" % len(names)) and "sm_100a" in out[0]
    assert out[1] == "rc 0 len 8"
    lo, hi = orc.wave_filters("db4")
    got = np.array([[float(v) for v in ln.split()] for ln in out[2:10]])
    assert np.array_equal(got[:, 0], lo) and np.array_equal(got[:, 1], hi)      # taps to the last bit (wave_filters.m:21-156)
    assert out[10] == "nb 46"                                                   # BASELINE configs[3]: 1 + 3 (2^4 - 1)
    assert out[11].startswith("rc -2") and "Unknown Wavelet Name" in out[11]


def test_m_files_call_the_gateway_the_way_it_is_written():
    """The .m classes cannot be executed here (no MATLAB), so at least their calls into nd_dwt_mex are checked against
    the gateway source: every string command they use exists there and is called with an argument count it accepts,
    the five-argument transform call has the reference's shape (mex/nd_dwt_mex.c:8-9), and the option keys handled by
    nddwt_b200_setup.m are the ones the Python mirror (api.py) handles."""
    mdir = os.path.join(ROOT, "non-decimated_wavelets_b200", "matlab")
    gw = open(os.path.join(mdir, "nd_dwt_mex.cpp")).read()
    accepts = {}
    for cmd, op, n in re.findall(r'strcmp\(cmd, "(\w+)"\) == 0(?: && nrhs (==|>=) (\d+))?', gw):
        accepts[cmd] = (op or ">=", int(n) if n else 1)
    assert set(accepts) == {"taps", "plan", "shrink", "dilations", "release"}
    calls = []
    for fn in sorted(os.listdir(mdir)):
        if not fn.endswith(".m"):
            continue
        for line in open(os.path.join(mdir, fn)):
            code = line.split("%")[0]
            for m in re.finditer(r"nd_dwt_mex\(([^;]*)\)", code):
                args = [a.strip() for a in re.split(r",(?![^\[\]]*\])", m.group(1).rstrip(" ,.)"))]
                calls.append((fn, args))
    assert len(calls) >= 6
    for fn, args in calls:
        if args[0].startswith("'"):
            cmd = args[0].strip("'")
            assert cmd in accepts, (fn, cmd)
            op, n = accepts[cmd]
            assert (len(args) == n) if op == "==" else (len(args) >= n), (fn, args)
        else:
            assert len(args) == 5, (fn, args)          # y = nd_dwt_mex(x, f, dir, level, pres_l2_norm)
    setup = open(os.path.join(mdir, "nddwt_b200_setup.m")).read()
    m_keys = set(re.findall(r"strcmp\(key, '(\w+)'\)", setup))
    api = open(os.path.join(ROOT, "non-decimated_wavelets_b200", "api.py")).read()
    for key in m_keys:
        assert '"%s"' % key in api, key
    assert {"pres_l2_norm", "compute", "precision"} <= m_keys
    for text in ("ingoring!", "of Data is shorter than the wavelet filter being used"):     # the reference's own texts
        assert text in setup and text in api
