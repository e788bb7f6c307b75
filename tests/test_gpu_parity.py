"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(ctypes -> libnddwt_b200.so); the oracle (numpy restatement of the reference) is only the checker.
Tolerances are BASELINE.json's: relative L2 <= 1e-5 single, <= 1e-12 double."""
import glob
import os

import numpy as np
import pytest

import nddwt_b200 as nd
from oracle import nddwt_oracle as orc
from conftest import TOL

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
CLS = {1: nd.nd_dwt_1D, 2: nd.nd_dwt_2D, 3: nd.nd_dwt_3D, 4: nd.nd_dwt_4D}


def _prec(dt):
    return "single" if np.dtype(dt) in (np.dtype(np.float32), np.dtype(np.complex64)) else "double"


def _obj(sizes, wn, l2, prec, compute="mex", kernel_mode=0):
    d = len(sizes)
    w = wn if (d > 1 or isinstance(wn, str)) else wn[0]
    o = CLS[d](w, list(sizes), "pres_l2_norm", int(l2), "precision", prec, "compute", compute)
    o.set_kernel_mode(kernel_mode)
    return o


@pytest.mark.parametrize("kernel_mode", [0, 1], ids=["auto", "generic"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(path, kernel_mode):
    z = np.load(path)
    wn = [str(s) for s in z["wname"]]
    wn = wn[0] if len(wn) == 1 else wn
    x, y, level, l2 = z["x"], z["y"], int(z["level"]), bool(z["pres_l2"])
    prec = _prec(x.dtype)
    o = _obj(x.shape, wn, l2, prec, kernel_mode=kernel_mode)
    yg = o.dec(x, level)
    assert yg.shape == y.shape and yg.dtype == y.dtype
    assert orc.rel_l2(yg, y) <= TOL[prec]
    xg = o.rec(y)
    assert xg.dtype == x.dtype
    assert orc.rel_l2(xg, x) <= TOL[prec]
    assert orc.rel_l2(o.rec(yg), x) <= TOL[prec]


# the reference's own test shapes (Test/nddwt{1,2,3,4}D_test.m:5-8, mex/mex_test.m:11,48,84,120,
# SURVEY.md 8c) -- odd, non-power-of-two sizes and mixed wavelets included
REF_SHAPES = [
    ((54321,), "db1", 4, 0),
    ((10000,), "db1", 3, 1),
    ((264, 264), ["db1", "db3"], 1, 1),
    ((129, 131), "db3", 2, 0),
    ((164, 64, 40), ["db1", "db3", "db1"], 1, 1),
    ((131, 128, 30), "db3", 2, 0),
    ((64, 64, 20), ["db1", "db3", "db9"], 2, 0),
    ((64, 64, 20, 10), ["db1", "db3", "db1", "db1"], 1, 1),
    ((32, 32, 16, 16), ["db1", "db3", "db3", "db5"], 2, 0),
    ((128, 68, 8, 8), "db3", 1, 0),
    ((256, 256), "db4", 3, 0),                      # BASELINE configs[0]
    ((256, 256), ["db1", "db4"], 2, 1),             # example_nd_dwt_2D.m:5-8 literal parameters
    ((64, 48, 40), "db4", 3, 0),
    ((32, 32, 24, 16), "db4", 3, 1),
    ((4096,), "db8", 6, 0),
    ((65536,), "db8", 6, 0),                        # one signal of BASELINE configs[1] (the reference has no batch API)
    ((40, 36, 20, 12), "db1", 3, 0),
]


@pytest.mark.parametrize("dtype", ["complex64", "complex128", "float32", "float64"])
@pytest.mark.parametrize("sizes,wn,level,l2", REF_SHAPES, ids=[("x".join(map(str, c[0])) + "_" + (c[1] if isinstance(c[1], str) else "mix") + "_J%d" % c[2]) for c in REF_SHAPES])
def test_parity_vs_oracle(sizes, wn, level, l2, dtype):
    prec = _prec(dtype)
    if np.prod(sizes) > 300000 and dtype in ("float32", "float64") and len(sizes) == 4:
        pytest.skip("covered by the complex cases")
    x = orc.synth(sizes, dtype, 11)
    o = _obj(sizes, wn, l2, prec)
    y = o.dec(x, level)
    yo = orc.dec_direct(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), wn, level, bool(l2))
    assert y.dtype == np.dtype(dtype)
    assert y.shape == yo.shape
    assert orc.rel_l2(y, yo) <= TOL[prec]
    xr = o.rec(y)
    assert orc.rel_l2(xr, x) <= TOL[prec]                                   # P1
    if l2:
        assert abs(np.linalg.norm(y.ravel()) / np.linalg.norm(x.ravel()) - 1) < 10 * TOL[prec]   # P2
    # rec of arbitrary (non-image) coefficients == adjoint per the oracle
    c = orc.synth(yo.shape, dtype, 12)
    xo = orc.rec_direct(c.astype(np.complex128 if np.iscomplexobj(c) else np.float64), wn, bool(l2))
    assert orc.rel_l2(o.rec(c), xo) <= TOL[prec]


@pytest.mark.parametrize("sizes,wn,level,fused", [
    ((164, 64, 40), ["db1", "db3", "db1"], 1, 1),              # Test/nddwt3D_test.m:5-7
    ((64, 64, 20, 10), ["db1", "db3", "db1", "db1"], 1, 1),    # Test/nddwt4D_test.m:5-7
    ((264, 264), ["db1", "db3"], 1, 1),                        # Test/nddwt2D_test.m:5-8
    ((256, 256), ["db1", "db4"], 2, 1),                        # example_nd_dwt_2D.m:5-8
    ((48, 40, 24), ["db4", "db2", "db3"], 2, 1),
    ((64, 64, 20), ["db1", "db3", "db9"], 2, 2),               # Test: db9 exceeds the register ring -> hybrid: generic pass along dim 3, fused 2-D kernels over the planes
    ((32, 32, 16, 16), ["db1", "db3", "db3", "db5"], 2, 2),    # SURVEY 8c shape: db5 along dim 4 -> hybrid
    ((40, 36, 24), "db7", 2, 2),                               # long filters everywhere -> hybrid
    ((131, 128, 30), "db3", 2, 1),                             # mex/mex_test.m:84: odd rows -> element-wise tile-kernel instantiations
    ((33, 24, 10, 8), "db2", 2, 1),                            # 4-D, odd rows, hyperplane still a 16-byte multiple
    ((33, 25, 9, 8), "db2", 1, 2),                             # 4-D, all-odd hyperplane: no 16-byte last-dim passes -> hybrid
])
def test_reference_shapes_kernel_family_and_parity(sizes, wn, level, fused):
    """The reference's own test shapes with MIXED wavelets run the fused kernels (the shorter filters are
    zero-padded to the longest tap length, which keeps their phase) and equal the generic kernels and the oracle.
    fused: 1 = fused tile kernels, 2 = hybrid (generic passes along the outer dims, fused 2-D kernels over the planes)."""
    x = orc.synth(sizes, np.complex64, 8)
    a = _obj(sizes, wn, 1, "single", kernel_mode=0)
    b = _obj(sizes, wn, 1, "single", kernel_mode=1)
    ya = a.dec(x, level)
    assert a._plan(True, 0).last_path == fused
    assert orc.rel_l2(ya, orc.dec_direct(x.astype(np.complex128), wn, level, True)) <= 1e-5
    assert orc.rel_l2(ya, b.dec(x, level)) <= 2e-6
    c = orc.synth(ya.shape, np.complex64, 9)
    xa = a.rec(c)
    assert a._plan(True, 0).last_path == fused
    assert orc.rel_l2(xa, orc.rec_direct(c.astype(np.complex128), wn, True)) <= 1e-5
    assert orc.rel_l2(a.rec(ya), x) <= 1e-5


@pytest.mark.parametrize("sizes,wn,level,l2", [((64, 48, 40), "db4", 3, 0), ((129, 131), "db3", 2, 1),
                                                ((32, 32, 24, 16), "db4", 2, 0), ((4099,), "db8", 4, 0)])
def test_fused_equals_generic(sizes, wn, level, l2):
    x = orc.synth(sizes, np.complex64, 5)
    a = _obj(sizes, wn, l2, "single", kernel_mode=0)
    b = _obj(sizes, wn, l2, "single", kernel_mode=1)
    ya, yb = a.dec(x, level), b.dec(x, level)
    assert orc.rel_l2(ya, yb) <= 2e-6
    assert orc.rel_l2(a.rec(ya), b.rec(ya)) <= 2e-6


@pytest.mark.parametrize("wn", ["db1", "db4", "db7", "db10"])
@pytest.mark.parametrize("dtype", ["complex64", "complex128", "float32", "float64"])
def test_fused_2d_kernels(wn, dtype):
    """2-D fused level kernels (any tap length, odd sizes) == generic kernels == oracle."""
    sizes = (131, 77)
    prec = _prec(dtype)
    x = orc.synth(sizes, dtype, 21)
    a = _obj(sizes, wn, 1, prec, kernel_mode=0)
    b = _obj(sizes, wn, 1, prec, kernel_mode=1)
    ya = a.dec(x, 3)
    assert a._plan(np.iscomplexobj(x), 0).last_path == 1 and b.dec(x, 1) is not None
    yo = orc.dec_direct(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), wn, 3, True)
    assert orc.rel_l2(ya, yo) <= TOL[prec]
    assert orc.rel_l2(ya, b.dec(x, 3)) <= 10 * TOL[prec]
    assert orc.rel_l2(a.rec(ya), x) <= TOL[prec]
    c = orc.synth(yo.shape, dtype, 22)
    assert orc.rel_l2(a.rec(c), b.rec(c)) <= 10 * TOL[prec]


def test_matches_fft_mat_path_and_mex_flow():
    """P3: same answer as the restated 'mat' FFT path and the MEX slot flow."""
    x = orc.synth((48, 40, 24), np.complex128, 2)
    o = _obj(x.shape, "db4", 0, "double")
    y = o.dec(x, 3)
    assert orc.rel_l2(y, orc.dec(x, "db4", 3)) <= 1e-12
    assert orc.rel_l2(y, orc.dec_mex(x, "db4", 3)) <= 1e-12


def test_haar_classes():
    for l2 in (0, 1):
        x = orc.synth((24, 18), np.complex128, 4)
        h = nd.harr_nddwt_2D([24, 18], "pres_l2_norm", l2)
        y = h.dec(x, 1)
        assert orc.rel_l2(y, orc.haar_level_1_dec(x, bool(l2))) <= 1e-12
        assert orc.rel_l2(h.rec(y), orc.haar_level_1_rec(y, bool(l2))) <= 1e-12
        x4 = orc.synth((10, 8, 6, 8), np.complex64, 5)
        h4 = nd.harr_nddwt_4D([10, 8, 6, 8], "pres_l2_norm", l2, "precision", "single")
        y4 = h4.dec(x4, 1)
        assert orc.rel_l2(y4, orc.haar_level_1_dec(x4.astype(np.complex128), bool(l2))) <= 1e-5
        assert orc.rel_l2(h4.rec(y4), x4) <= 1e-5


def test_device_resident_path_and_input_not_mutated():
    import torch
    x = orc.synth((40, 36, 20), np.complex64, 9)
    o = nd.nd_dwt_3D("db4", [40, 36, 20], "precision", "single", "compute", "gpu")
    xd = nd.to_device(x)
    yd = o.dec(xd, 2)
    assert isinstance(yd, torch.Tensor) and yd.is_cuda and tuple(yd.shape) == (40, 36, 20, 15)
    y = nd.to_host(yd)
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), "db4", 2)) <= 1e-5
    ykeep = yd.clone()
    xr = o.rec(yd)
    assert torch.equal(yd, ykeep)          # unlike nddwt.c:163 the coefficients are not modified
    assert orc.rel_l2(nd.to_host(xr), x) <= 1e-5
    assert torch.equal(xd, nd.to_device(x))


@pytest.mark.parametrize("sizes,wn,level", [((256, 256), "db4", 3), ((4096,), "db8", 4), ((64, 48, 40), "db4", 2), ((32, 32, 24, 16), "db4", 2)])
def test_cuda_graph_capture_and_replay(sizes, wn, level):
    """Once a plan's scratch exists, dec and rec only launch kernels on the caller's stream (no allocation, no
    synchronisation, no host round trip), so the dec/rec pair of an iterative loop can be captured into a CUDA graph and
    replayed on new data -- what a launch-bound small problem (BASELINE configs[0], 256 x 256) wants."""
    import torch
    x1 = orc.synth(sizes, np.complex64, 77)
    x2 = orc.synth(sizes, np.complex64, 78)
    o = CLS[len(sizes)](wn, list(sizes), "precision", "single", "compute", "gpu")
    xd = nd.to_device(x1)
    o.rec(o.dec(xd, level))                      # warm-up outside the capture: plan scratch gets allocated here
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        yd = o.dec(xd, level)
        xr = o.rec(yd)
    xd.copy_(nd.to_device(x2))                   # new input in the captured buffer
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    assert orc.rel_l2(nd.to_host(yd), orc.dec_direct(x2.astype(np.complex128), wn, level)) <= 1e-5
    assert orc.rel_l2(nd.to_host(xr), x2) <= 1e-5


@pytest.mark.parametrize("sizes,wn,levels", [((96, 80), ["db2", "db4"], (3, 2, 4)), ((48, 40, 24), "db4", (2, 3)),
                                             ((32, 24, 16, 12), "db4", (3, 2)), ((40, 33, 20), ["db1", "db3", "db9"], (2, 3))])
@pytest.mark.parametrize("dtype", ["complex64", "float64"])
def test_host_entry_points_stream_levels(sizes, wn, levels, dtype):
    """nddwt_dec_host / nddwt_rec_host on 2-D ... 4-D arrays with two or more levels hold two LEVEL buffers on the device
    (not the whole coefficient stack) and move the bands of one level over PCIe while the next level computes: the
    result equals the device-pointer path and the oracle, across repeated calls and changing level counts on one plan."""
    prec = _prec(dtype)
    x = orc.synth(sizes, dtype, 91)
    host = _obj(sizes, wn, 0, prec, compute="mex")
    dev = _obj(sizes, wn, 0, prec, compute="gpu")
    for level in levels:
        for rep in range(2):
            y = host.dec(x, level)
            yd = nd.to_host(dev.dec(nd.to_device(x), level))
            assert np.array_equal(y, yd)                      # same kernels, same order: bit-identical
            yo = orc.dec_direct(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), wn, level)
            assert orc.rel_l2(y, yo) <= TOL[prec]
            c = orc.synth(y.shape, dtype, 92 + rep)
            assert np.array_equal(host.rec(c), nd.to_host(dev.rec(nd.to_device(c))))
            assert orc.rel_l2(host.rec(y), x) <= TOL[prec]


def test_nd_dwt_mex_entry():
    x = orc.synth((32, 20), np.complex128, 1)
    f = nd.FilterSpec(["db2", "db3"], [32, 20])
    y = nd.nd_dwt_mex(x, f, 0, 2, 1)
    assert y.shape == (32, 20, 7)
    assert orc.rel_l2(y, orc.dec(x, ["db2", "db3"], 2, True)) <= 1e-12
    assert orc.rel_l2(nd.nd_dwt_mex(y, f, 1, 2, 1), x) <= 1e-12


def test_linearity_and_shift_equivariance_full_size():
    """Size-independent properties at a BASELINE size (cfg3: 256^3 complex single, db4, 3 levels)."""
    import torch
    n = 256
    o = nd.nd_dwt_3D("db4", [n, n, n], "precision", "single", "compute", "gpu")
    g = torch.Generator(device="cuda").manual_seed(0)
    def rnd():
        t = torch.randn((n, n, n, 2), generator=g, device="cuda", dtype=torch.float32)
        return torch.view_as_complex(t).permute(2, 1, 0)
    a, b = rnd(), rnd()
    ya, yb = o.dec(a, 3), o.dec(b, 3)
    yab = o.dec(a + 2 * b, 3)
    num = torch.linalg.vector_norm(yab - (ya + 2 * yb)); den = torch.linalg.vector_norm(yab)
    assert float(num / den) <= 1e-5
    del yab, yb
    # circular shift equivariance: dec(roll(a)) == roll(dec(a))
    ys = o.dec(torch.roll(a, shifts=(3, 5, 7), dims=(0, 1, 2)), 3)
    num = torch.linalg.vector_norm(ys - torch.roll(ya, shifts=(3, 5, 7), dims=(0, 1, 2)))
    assert float(num / torch.linalg.vector_norm(ys)) <= 1e-5
    del ys
    xr = o.rec(ya)
    assert float(torch.linalg.vector_norm(xr - a) / torch.linalg.vector_norm(a)) <= 1e-5
    # energy: frame bound 2^d per level without pres_l2_norm is not Parseval; use pres_l2_norm object
    o2 = nd.nd_dwt_3D("db4", [n, n, n], "precision", "single", "compute", "gpu", "pres_l2_norm", 1)
    y2 = o2.dec(a, 3)
    assert abs(float(torch.linalg.vector_norm(y2) / torch.linalg.vector_norm(a)) - 1) <= 1e-4


@pytest.mark.parametrize("dtype", ["float32", "float64", "complex64", "complex128"])
@pytest.mark.parametrize("sizes,wn,level", [((131, 70, 30), "db4", 2), ((67, 41, 12, 6), "db3", 2), ((35, 33, 20), ["db1", "db4", "db2"], 2)])
def test_odd_row_lengths_fused(sizes, wn, level, dtype):
    """Rows that are not 16-byte multiples in 3-D / 4-D: element-wise instantiations of the tile kernels, every dtype
    (complex double rows are always 16-byte multiples and take the normal path)."""
    prec = _prec(dtype)
    x = orc.synth(sizes, dtype, 14)
    a = _obj(sizes, wn, 0, prec)
    wide = np.complex128 if np.iscomplexobj(x) else np.float64
    y = a.dec(x, level)
    assert a._plan(np.iscomplexobj(x), 0).last_path == 1
    assert orc.rel_l2(y, orc.dec_direct(x.astype(wide), wn, level)) <= TOL[prec]
    c = orc.synth(y.shape, dtype, 15)
    assert orc.rel_l2(a.rec(c), orc.rec_direct(c.astype(wide), wn, False)) <= TOL[prec]
    assert a._plan(np.iscomplexobj(x), 0).last_path == 1
    assert orc.rel_l2(a.rec(y), x) <= TOL[prec]
    # the fused shrink has no element-wise instantiation: such plans fall back to the generic kernels + in-place pass
    if dtype == "complex64":
        tab = np.full((level, 1 << len(sizes)), 0.3)
        a.set_shrink(tab)
        ys = a.dec(x, level)
        assert orc.rel_l2(ys, orc.shrink_soft(orc.dec_direct(x.astype(wide), wn, level), tab, len(sizes))) <= TOL[prec]


def test_cfg3_full_size_vs_oracle():
    """BASELINE configs[2] at full size (256^3 complex single, db4, 3 levels) against the oracle itself, not only
    through properties: 22 bands of 16.8 M voxels, compared band by band (the oracle needs ~40 s of CPU)."""
    n = 256
    x = orc.synth((n, n, n), np.complex64, 13)
    o = nd.nd_dwt_3D("db4", [n, n, n], "precision", "single", "compute", "gpu")
    yd = o.dec(nd.to_device(x), 3)
    assert o._plan(True, 0).last_path == 1
    xr = nd.to_host(o.rec(yd))
    assert orc.rel_l2(xr, x) <= 1e-5
    y = nd.to_host(yd)
    del yd
    yo = orc.dec_direct(x, "db4", 3)          # numpy complex64 arithmetic (rounding ~1e-7, two decades under the tolerance)
    assert yo.shape == y.shape
    worst = 0.0
    for b in range(y.shape[-1]):
        worst = max(worst, orc.rel_l2(y[..., b], yo[..., b]))
    assert worst <= 1e-5, worst


@pytest.mark.parametrize("sizes,level", [((256, 24, 10, 16), 2), ((256, 40, 9, 8), 1), ((256, 17, 12, 8), 1), ((256, 32, 16, 12), 3)])
def test_cfg4_row_length_4d(sizes, level):
    """Rows of 256 elements (cfg4's row length): the 512-thread full-row synthesis instantiation and the
    analysis tile kernel on several 4-D shapes (partial last row block, dim-2 wrap inside a tile, short dims)."""
    x = orc.synth(sizes, np.complex64, 23)
    a = _obj(sizes, "db4", 0, "single")
    a.set_param("rows_min_ctas", 0)
    g = _obj(sizes, "db4", 0, "single", kernel_mode=1)
    y = a.dec(x, level)
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), "db4", level)) <= 1e-5
    c = orc.synth(y.shape, np.complex64, 24)
    xa = a.rec(c)
    assert a.synthesis_kernels() == [4]
    assert orc.rel_l2(xa, orc.rec_direct(c.astype(np.complex128), "db4", False)) <= 1e-5
    assert orc.rel_l2(xa, g.rec(c)) <= 2e-6
    assert orc.rel_l2(a.rec(y), x) <= 1e-5


def test_cfg2_batch_of_signals():
    """BASELINE configs[1] at full size (4096 signals of 65536 samples, complex single, db8, 6 levels: 2.1 GB in,
    15 GB of coefficients): the batched cascade equals the oracle on sampled signals and reconstructs every signal."""
    import torch
    if torch.cuda.mem_get_info()[0] < 30e9:
        pytest.skip("needs ~22 GB of device memory")
    n, B, level = 65536, 4096, 6
    o = nd.nd_dwt_1D("db8", n, "precision", "single", "compute", "gpu")
    g = torch.Generator(device="cuda").manual_seed(3)
    xb = torch.view_as_complex(torch.randn((B, n, 2), generator=g, device="cuda", dtype=torch.float32))   # [B][n] = column-major [n, B]
    x = xb.permute(1, 0)
    y = o.dec(x, level)
    assert tuple(y.shape) == (n, B, level + 1)
    xr = o.rec(y)
    assert float(torch.linalg.vector_norm(xr - x) / torch.linalg.vector_norm(x)) <= 1e-5
    for b in (0, 101, 2048, B - 1):
        yo = orc.dec_direct(x[:, b].cpu().numpy().astype(np.complex128), "db8", level)
        assert orc.rel_l2(y[:, b, :].cpu().numpy(), yo) <= 1e-5


def test_dilated_atrous_mode_opt_in():
    x = orc.synth((64, 48), np.complex128, 3)
    o = nd.nd_dwt_2D("db2", [64, 48])
    o.set_dilations([1, 2, 4])
    y = o.dec(x, 3)
    assert orc.rel_l2(y, orc.dec_direct(x, "db2", 3, dilations=[1, 2, 4])) <= 1e-12
    assert orc.rel_l2(o.rec(y), x) <= 1e-12


def test_atrous_option_and_haar_multilevel():
    """'atrous' option (extension): dilation 2^(j-1) at level j -- the transform the literature calls
    non-decimated -- for the Daubechies classes and as the multi-level Haar that harr_nddwt_4D.m:175,222 does not
    compute correctly.  Oracle: the closed-form restatement with the same dilations; perfect reconstruction."""
    x = orc.synth((24, 20, 16, 16), np.complex64, 6)
    h = nd.harr_nddwt_4D([24, 20, 16, 16], "precision", "single", "atrous", 1)
    y = h.dec(x, 3)
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), "db1", 3, dilations=[1, 2, 4])) <= 1e-5
    assert orc.rel_l2(h.rec(y), x) <= 1e-5
    x3 = orc.synth((40, 36, 32), np.complex128, 7)
    o = nd.nd_dwt_3D("db3", [40, 36, 32], "pres_l2_norm", 1, "atrous", True)
    y3 = o.dec(x3, 2)
    assert orc.rel_l2(y3, orc.dec_direct(x3, "db3", 2, True, dilations=[1, 2])) <= 1e-12
    assert orc.rel_l2(o.rec(y3), x3) <= 1e-12
    assert abs(np.linalg.norm(y3.ravel()) / np.linalg.norm(x3.ravel()) - 1) < 1e-11      # still a tight frame


@pytest.mark.parametrize("sizes,wn,dil,dtype,paths", [
    ((24, 20, 16, 16), "db1", [1, 2, 4], "complex64", [1, 1, 1]),      # a-trous Haar, 3 levels: 2-, 4-, 8-tap tile kernels
    ((40, 36, 32), "db1", [1, 2, 4], "complex64", [1, 1, 1]),
    ((40, 36, 32), "db2", [1, 2, 4], "float64", [1, 1, 2]),            # 16 stretched taps: dim 3 generic + fused 2-D planes
    ((32, 24, 16, 16), "db2", [1, 2], "complex64", [1, 1]),
    ((40, 36, 32), "db3", [1, 2], "complex128", [1, 2]),               # 12 stretched taps: hybrid
    ((96, 80), "db4", [1, 2], "complex64", [1, 1]),                    # 2-D: 16 stretched taps
    ((96, 80), "db5", [1, 2], "float32", [1, 1]),                      # 20 taps, the longest instantiation
    ((96, 80), "db1", [1, 2, 4, 8], "complex128", [1, 1, 1, 1]),
    ((96, 80), "db4", [1, 2, 4], "complex64", [1, 1, 0]),              # 32 taps: generic
    ((131, 30, 20), ["db1", "db2", "db1"], [1, 2], "complex64", [1, 1]),   # mixed wavelets, odd rows, stretched
])
def test_atrous_levels_take_the_fused_kernels(sizes, wn, dil, dtype, paths):
    """A-trous levels run the SAME fused kernels with their taps stretched about the centre (an L-tap
    filter at dilation s is an L s-tap filter with zeros in between): Haar at dilations 1, 2, 4 -- the multi-level Haar the
    reference's harr_nddwt_4D.m:175,222 cannot compute -- takes the 2-, 4- and 8-tap tile kernels; the 2-D kernels go up to
    20 stretched taps; beyond that the level is hybrid (generic outer passes + fused 2-D planes) or generic.
    Oracle: the closed-form restatement with the same dilations; kernel family per level from the plan."""
    prec = _prec(dtype)
    x = orc.synth(sizes, dtype, 61)
    cplx = np.iscomplexobj(x)
    a = _obj(sizes, wn, 1, prec, kernel_mode=0)
    b = _obj(sizes, wn, 1, prec, kernel_mode=1)
    a.set_dilations(dil)
    b.set_dilations(dil)
    level = len(dil)
    ya = a.dec(x, level)
    yo = orc.dec_direct(x.astype(np.complex128 if cplx else np.float64), wn, level, True, dilations=dil)
    assert orc.rel_l2(ya, yo) <= TOL[prec]
    assert orc.rel_l2(ya, b.dec(x, level)) <= 10 * TOL[prec]
    assert orc.rel_l2(a.rec(ya), x) <= TOL[prec]
    c = orc.synth(ya.shape, dtype, 62)
    assert orc.rel_l2(a.rec(c), b.rec(c)) <= 10 * TOL[prec]             # adjoint on arbitrary coefficients
    assert abs(np.linalg.norm(ya.ravel()) / np.linalg.norm(x.ravel()) - 1) < 1e3 * TOL[prec]   # still a tight frame
    # kernel family of each level: run them one at a time (a one-level transform with that level's dilation)
    pa = a._plan(cplx, 0)
    for s_, want in zip(dil, paths):
        a.set_dilations([s_])
        a.dec(x, 1)
        assert pa.last_path == want, (s_, pa.last_path, want)
        y1 = a.dec(x, 1)
        a.rec(y1)
        assert pa.last_path == want, ("rec", s_, pa.last_path, want)


def test_4d_full_size_properties_cfg5():
    """Size-independent properties at BASELINE configs[4]'s full size (192x192x64x48 complex single,
    db4, 3 levels; the bench.py N=1 workload): perfect reconstruction, linearity, energy."""
    import torch
    sizes = [192, 192, 64, 48]
    if torch.cuda.mem_get_info()[0] < 120e9:
        pytest.skip("needs ~100 GB of device memory")
    o = nd.nd_dwt_4D("db4", sizes, "precision", "single", "compute", "gpu", "pres_l2_norm", 1)
    g = torch.Generator(device="cuda").manual_seed(1)
    def rnd():
        t = torch.randn(tuple(reversed(sizes)) + (2,), generator=g, device="cuda", dtype=torch.float32)
        return torch.view_as_complex(t).permute(3, 2, 1, 0)
    a = rnd()
    ya = o.dec(a, 3)
    assert tuple(ya.shape) == tuple(sizes) + (46,)
    na = float(torch.linalg.vector_norm(a))
    assert abs(float(torch.linalg.vector_norm(ya)) / na - 1) <= 1e-4          # P2: Parseval with pres_l2_norm
    xr = o.rec(ya)
    assert float(torch.linalg.vector_norm(xr - a)) / na <= 1e-5               # P1
    del xr
    b = rnd()
    yab = o.dec(a + 0.5 * b, 3)
    yab -= ya
    del ya
    yb = o.dec(b, 3)
    yab -= 0.5 * yb
    assert float(torch.linalg.vector_norm(yab)) / float(torch.linalg.vector_norm(yb)) <= 1e-5   # linearity


@pytest.mark.parametrize("sizes,wn,level,batch", [((4096,), "db8", 6, 8), ((1001,), "db5", 4, 3), ((64, 48), ["db2", "db3"], 2, 5),
                                                    ((32, 24, 16), "db4", 2, 3)])
def test_batched_extension(sizes, wn, level, batch):
    """Batch API (extension): x is [sizes, B]; every slice equals the un-batched transform."""
    x = orc.synth(tuple(sizes) + (batch,), np.complex64, 31)
    o = _obj(sizes, wn, 0, "single")
    y = o.dec(x, level)
    nb = orc.num_bands(len(sizes), level)
    assert y.shape == tuple(sizes) + (batch, nb)
    for b in (0, batch - 1):
        yo = orc.dec_direct(x[..., b].astype(np.complex128), wn, level)
        assert orc.rel_l2(y[..., b, :], yo) <= 1e-5
    assert orc.rel_l2(o.rec(y), x) <= 1e-5


@pytest.mark.parametrize("n,wn,level", [(54321, "db1", 4), (4099, "db8", 6), (300, "db10", 3), (61, "db3", 5), (65536, "db4", 8),
                                        (16, "db8", 3), (18, "db7", 2), (7000, "db2", 5), (2, "db1", 3), (3586, "db8", 6), (12288, "db6", 1)])
@pytest.mark.parametrize("dtype", ["complex64", "float64", "float32", "complex128"])
def test_fused_1d_cascade(n, wn, level, dtype):
    """1-D cascade kernel (all levels in one launch) == generic per-level kernels == oracle."""
    prec = _prec(dtype)
    x = orc.synth((n,), dtype, 41)
    a = _obj((n,), wn, 1, prec, kernel_mode=0)
    b = _obj((n,), wn, 1, prec, kernel_mode=1)
    ya = a.dec(x, level)
    pa = a._plan(np.iscomplexobj(x), 0)
    l0 = pa.launches
    ya = a.dec(x, level)
    assert pa.launches - l0 == 1 and pa.last_path == 1
    yo = orc.dec_direct(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), wn, level, True)
    assert orc.rel_l2(ya, yo) <= TOL[prec]
    assert orc.rel_l2(ya, b.dec(x, level)) <= 10 * TOL[prec]
    assert orc.rel_l2(a.rec(ya), x) <= TOL[prec]
    c = orc.synth(yo.shape, dtype, 42)
    assert orc.rel_l2(a.rec(c), b.rec(c)) <= 10 * TOL[prec]


@pytest.mark.parametrize("wn,dtype", [("db1", "complex64"), ("db2", "complex64"), ("db3", "complex64"), ("db4", "complex64"),
                                      ("db2", "float64"), ("db4", "float64")])
@pytest.mark.parametrize("sizes,level", [((64, 30, 9, 8), 1), ((192, 40, 16, 8), 1), ((72, 20, 8, 8), 1), ((224, 24, 9, 8), 1),
                                          ((128, 17, 12), 1), ((32, 32, 24, 16), 2), ((64, 32, 12, 16), 2), ((256, 20, 8, 8), 1)])
def test_full_row_synthesis_kernel(sizes, level, wn, dtype):
    """k_rec3_rows (full-row tiles, 8-byte elements) is chosen by default only for big 4-D batches; lower
    its CTA bound so that small, odd shapes (dim-2 wrap inside a tile, partial last tile, dim-3/4 shorter
    than two rings) reach it, and compare with the generic kernels and the oracle."""
    prec = _prec(dtype)
    x = orc.synth(sizes, dtype, 21)
    a = _obj(sizes, wn, 0, prec, kernel_mode=0)
    b = _obj(sizes, wn, 1, prec, kernel_mode=0)
    g = _obj(sizes, wn, 0, prec, kernel_mode=1)
    a.set_param("rows_min_ctas", 0)
    b.set_param("rows_min_ctas", 0)
    y = a.dec(x, level)
    assert orc.rel_l2(a.rec(y), x) <= TOL[prec]
    assert a.synthesis_kernels() == [4], a.synthesis_kernels()
    wide = np.complex128 if np.iscomplexobj(x) else np.float64
    c = orc.synth(y.shape, dtype, 22)
    assert orc.rel_l2(a.rec(c), orc.rec_direct(c.astype(wide), wn, False)) <= TOL[prec]
    assert orc.rel_l2(b.rec(c), orc.rec_direct(c.astype(wide), wn, True)) <= TOL[prec]
    assert orc.rel_l2(a.rec(c), g.rec(c)) <= (2e-6 if prec == "single" else 1e-13)



def test_plain_c_caller_dec_rec(tmp_path):
    """A C99 program (no C++, no Python, no torch) drives the library through include/nddwt_b200.h: plan, host-pointer dec
    and rec of a 3-D complex-single array, perfect reconstruction, the band count, and an error return with its text --
    what a C host application (or another language's FFI) sees."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "c_caller.c"
    src.write_text(r'''
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "nddwt_b200.h"
int main(void)
{
    const int64_t dims[3] = {48, 40, 24};
    const char *w[3] = {"db4", "db2", "db4"};
    const int level = 2;
    const int64_t n = dims[0] * dims[1] * dims[2], nb = nddwt_num_bands(3, level);
    float *x = malloc(sizeof(float) * 2 * n), *xr = malloc(sizeof(float) * 2 * n), *y = malloc(sizeof(float) * 2 * n * nb);
    nddwt_plan *p = NULL;
    double num = 0.0, den = 0.0, e2 = 0.0;
    int64_t i;
    unsigned s = 12345u;
    int rc;
    for (i = 0; i < 2 * n; ++i) { s = s * 1664525u + 1013904223u; x[i] = (float)((double)(s >> 8) / 8388608.0 - 1.0); }
    rc = nddwt_plan_create(&p, 3, dims, w, NDDWT_C64, 1, 0);
    if (rc) { printf("plan rc %d %s\n", rc, nddwt_last_error()); return 1; }
    rc = nddwt_dec_host(p, x, y, level);
    if (rc) { printf("dec rc %d %s\n", rc, nddwt_last_error()); return 1; }
    rc = nddwt_rec_host(p, y, xr, level);
    if (rc) { printf("rec rc %d %s\n", rc, nddwt_last_error()); return 1; }
    for (i = 0; i < 2 * n; ++i) { num += (double)(xr[i] - x[i]) * (xr[i] - x[i]); den += (double)x[i] * x[i]; }
    for (i = 0; i < 2 * n * nb; ++i) e2 += (double)y[i] * y[i];
    printf("bands %d pr %.3e energy %.6f\n", (int)nb, sqrt(num / den), sqrt(e2 / den));
    rc = nddwt_dec_host(p, x, y, 99);
    printf("bad level rc %d msg %s\n", rc, nddwt_last_error());
    nddwt_plan_destroy(p);
    free(x); free(xr); free(y);
    return 0;
}
''')
    libdir = os.path.dirname(nd.LIB_PATH)
    exe = tmp_path / "c_caller"
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lnddwt_b200", "-lm", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    f = out[0].split()
    assert f[0] == "bands" and int(f[1]) == 15 and float(f[3]) <= 1e-5 and abs(float(f[5]) - 1.0) <= 1e-4     # pres_l2_norm: tight frame
    assert out[1].startswith("bad level rc -1") and "level" in out[1]
