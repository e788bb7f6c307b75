"""GPU tests of the fused coefficient-domain shrink (nddwt_plan_set_shrink): the soft threshold
applied in the analysis kernels' epilogue equals the oracle's shrink of the oracle's dec, on every kernel
family (1-D cascade, 2-D fused, 3-D / 4-D tile kernels, generic separable kernels, the multi-GPU plan)."""
import importlib

import numpy as np
import pytest

import nddwt_b200 as nd
from oracle import nddwt_oracle as orc
from conftest import TOL

pytestmark = pytest.mark.gpu
CLS = {1: nd.nd_dwt_1D, 2: nd.nd_dwt_2D, 3: nd.nd_dwt_3D, 4: nd.nd_dwt_4D}


def _table(d, level, seed, scale):
    rng = np.random.default_rng(seed)
    t = scale * (0.2 + rng.random((level, 1 << d)))
    t[:, 0] = 123.0          # column 0 must be ignored: the approximation band is never thresholded
    return t


CASES = [
    ((5000,), "db4", 4), ((300,), "db10", 3),
    ((131, 77), "db3", 3), ((64, 48), ["db2", "db3"], 2),
    ((64, 48, 40), "db4", 3), ((32, 20, 12), ["db1", "db3", "db2"], 2),
    ((32, 32, 24, 16), "db4", 2), ((24, 20, 12, 8), "db2", 3), ((40, 36, 10, 8), "db1", 2),
]


@pytest.mark.parametrize("dtype", ["complex64", "complex128", "float32", "float64"])
@pytest.mark.parametrize("kernel_mode", [0, 1], ids=["auto", "generic"])
@pytest.mark.parametrize("sizes,wn,level", CASES, ids=["x".join(map(str, c[0])) for c in CASES])
def test_fused_shrink_matches_oracle(sizes, wn, level, kernel_mode, dtype):
    d = len(sizes)
    prec = "single" if dtype in ("complex64", "float32") else "double"
    x = orc.synth(sizes, dtype, 17)
    w = wn if (d > 1 or isinstance(wn, str)) else wn[0]
    o = CLS[d](w, list(sizes), "pres_l2_norm", 1, "precision", prec, "compute", "mex")
    o.set_kernel_mode(kernel_mode)
    wide = np.complex128 if np.iscomplexobj(x) else np.float64
    yo = orc.dec_direct(x.astype(wide), wn, level, True)
    table = _table(d, level, 5, 0.5 * float(np.sqrt(np.mean(np.abs(yo[..., -1]) ** 2))))
    o.set_shrink(table)
    y = o.dec(x, level)
    ref = orc.shrink_soft(yo, table, d)
    assert y.dtype == np.dtype(dtype)
    assert orc.rel_l2(y, ref) <= TOL[prec]
    assert np.array_equal(y[..., 0] == 0, ref[..., 0] == 0)
    frac = float(np.mean(ref[..., 1:] == 0))
    assert 0.05 < frac < 0.95, frac          # the thresholds bite, and not everything is killed
    # rec is untouched by the shrink state; switching it off restores the plain transform
    assert orc.rel_l2(o.rec(y), orc.rec_direct(ref, wn, True)) <= 10 * TOL[prec]
    o.set_shrink(None)
    assert orc.rel_l2(o.dec(x, level), yo) <= TOL[prec]


def test_scalar_and_per_level_thresholds_and_unfused_form():
    import torch
    sizes, level = (32, 32, 24, 16), 3
    x = orc.synth(sizes, np.complex64, 3)
    o = nd.nd_dwt_4D("db4", list(sizes), "precision", "single", "compute", "gpu")
    yo = orc.dec_direct(x.astype(np.complex128), "db4", level)
    o.set_shrink(0.7)
    y = nd.to_host(o.dec(nd.to_device(x), level))
    assert orc.rel_l2(y, orc.shrink_soft(yo, np.full((level, 16), 0.7), 4)) <= 1e-5
    o.set_shrink([0.2, 0.5, 1.1])
    yd = o.dec(nd.to_device(x), level)
    tab = np.repeat(np.array([0.2, 0.5, 1.1])[:, None], 16, axis=1)
    ref = orc.shrink_soft(yo, tab, 4)
    assert orc.rel_l2(nd.to_host(yd), ref) <= 1e-5
    # unfused form: plain dec, then nddwt_shrink in place on the stack
    plan = o._plan(True, 0)
    plan.set_shrink(None)
    y2 = o.dec(nd.to_device(x), level)
    base = y2.permute(*reversed(range(y2.dim())))
    assert base.is_contiguous()
    plan.set_shrink(tab)
    plan.shrink(base.data_ptr(), level, torch.cuda.current_stream().cuda_stream)
    assert orc.rel_l2(nd.to_host(y2), ref) <= 1e-5
    with pytest.raises(ValueError):
        o.set_shrink(-1.0)


def test_shrink_through_the_multi_gpu_plan():
    """Emulated ranks on one GPU (chunked, overlapped schedule): thresholds reach every rank's kernels."""
    import torch
    slab = importlib.import_module("non-decimated_wavelets_b200.slab")
    _lib = importlib.import_module("non-decimated_wavelets_b200._lib")
    sizes, wname, level = (32, 16, 40, 12), "db4", 2
    x = orc.synth(sizes, np.complex64, 9)
    tr = slab.MultiGpuTransform(sizes, [wname] * 4, _lib.NDDWT_C64, 0, [0, 0, 0])
    tab = _table(4, level, 2, 1.0)
    tr.plan.set_shrink(tab)
    xs = tr.scatter_input(nd.to_device(x).permute(3, 2, 1, 0).contiguous())
    ys = tr.dec(xs, level)
    y = nd.to_host(torch.cat(ys, dim=1).permute(*reversed(range(5))))
    assert orc.rel_l2(y, orc.shrink_soft(orc.dec_direct(x.astype(np.complex128), wname, level), tab, 4)) <= 1e-5
