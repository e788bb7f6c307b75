"""GPU tests of the multi-GPU plan (nddwt_mplan_*, csrc/nddwt_multi.cu) through the C ABI.

* On any box: the ranks are emulated on ONE GPU by one process (the same device listed several times), so
  the whole schedule -- routing of multi-hop halos, peer pushes, overlapped part-wise analysis,
  scatter-form synthesis with stage buffers -- is compared with the oracle, not only the kernels.
* With >= 2 GPUs visible: the same plan over real devices (one process), and the one-process-per-GPU form
  (CUDA IPC + flags in peer memory) through torchrun (tests/multi_gpu_check.py).
Tolerance: BASELINE.json's relative L2 <= 1e-5 (complex single)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

import nddwt_b200 as nd
from oracle import nddwt_oracle as orc

pytestmark = pytest.mark.gpu
slab = importlib.import_module("non-decimated_wavelets_b200.slab")
_lib = importlib.import_module("non-decimated_wavelets_b200._lib")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(sizes, wname, level, l2, devices, dil=None, kernel_mode=0, seed=3, params=None):
    import torch
    d = len(sizes)
    wn = [wname] * d if isinstance(wname, str) else list(wname)
    x = orc.synth(sizes, np.complex64, seed)
    xbase = nd.to_device(x).permute(*reversed(range(d))).contiguous()        # [N_d, ..., N_1]
    tr = slab.MultiGpuTransform(sizes, wn, _lib.NDDWT_C64, l2, devices, dilations=dil, kernel_mode=kernel_mode)
    for k, v in (params or {}).items():
        tr.plan.set_param(k, v)
    xs = tr.scatter_input(xbase)
    ys = tr.dec(xs, level)
    xr = tr.rec(ys)
    y = torch.cat([t.to("cuda:0") for t in ys], dim=1)                       # [nb, N_d, ..., N_1]
    xr = torch.cat([t.to("cuda:0") for t in xr], dim=0)
    y_np = nd.to_host(y.permute(*reversed(range(d + 1))))
    xr_np = nd.to_host(xr.permute(*reversed(range(d))))
    # synthesis of arbitrary coefficients (adjoint), not only of an image of dec
    c = orc.synth(y_np.shape, np.complex64, seed + 1)
    cbase = nd.to_device(c).permute(*reversed(range(d + 1))).contiguous()     # [nb, N_d, ..., N_1]
    cs = [cbase[:, s:s + n].to("cuda:%d" % dev).contiguous() for (s, n), dev in zip(tr.parts, tr.devices)]
    xc = torch.cat([t.to("cuda:0") for t in tr.rec(cs)], dim=0)
    xc_np = nd.to_host(xc.permute(*reversed(range(d))))
    assert tr.plan.wait_timeouts == 0
    return x, y_np, xr_np, c, xc_np, tr


CASES = [
    # sizes, wavelet, level, pres_l2, ranks
    ((32, 24, 16, 16), "db4", 3, 0, 4),        # 4 planes per rank, halo 3+4 wider than the slab (cfg4 at 8 GPUs in small)
    ((32, 24, 16, 12), "db4", 3, 1, 2),
    ((32, 20, 12, 7), "db2", 2, 0, 3),         # ragged slabs 3/2/2, halos reach two ranks and wrap to the rank itself
    ((40, 36, 10, 8), "db1", 2, 0, 4),         # Haar: one-sided halos
    ((64, 48, 24), "db4", 2, 0, 3),            # 3-D slabs: fused analysis, gather-form synthesis
    ((30, 20, 9), ["db1", "db3", "db2"], 2, 1, 2),   # mixed wavelets -> generic kernels
    ((48, 40), "db3", 2, 0, 3),                # 2-D slabs
    ((4000,), "db4", 3, 1, 4),                 # 1-D slabs
    ((32, 24, 16, 8), "db4", 2, 0, 8),         # one plane per rank: every halo plane from a different rank
    ((32, 16, 40, 12), "db4", 3, 0, 3),        # dim 3 long enough for 4 z-chunks: the pipelined (chunked) exchange
    ((32, 16, 36, 9), "db3", 2, 1, 4),         # chunked, ragged slabs 3/2/2/2, halo wider than the slabs
]


@pytest.mark.parametrize("sizes,wname,level,l2,world", CASES,
                         ids=["x".join(map(str, c[0])) + "_P%d" % c[4] for c in CASES])
def test_emulated_ranks_whole_schedule_vs_oracle(sizes, wname, level, l2, world):
    x, y, xr, c, xc, tr = _run(sizes, wname, level, l2, [0] * world)
    yo = orc.dec_direct(x.astype(np.complex128), wname, level, bool(l2))
    assert orc.rel_l2(y, yo) <= 1e-5
    assert orc.rel_l2(xr, x) <= 1e-5
    assert orc.rel_l2(xc, orc.rec_direct(c.astype(np.complex128), wname, bool(l2))) <= 1e-5
    if len(sizes) == 4 and isinstance(wname, str):
        assert tr.plan.separable            # the overlapped scatter schedule was the one tested


@pytest.mark.parametrize("zc", [1, 3, 5])
def test_emulated_ranks_chunk_counts(zc):
    """The z-chunk count of the pipelined exchange is a free parameter: uneven chunk boundaries (40 planes in 3 or
    5 chunks) and the unchunked schedule give the same answer."""
    sizes, wname, level = (32, 16, 40, 8), "db4", 2
    x, y, xr, c, xc, tr = _run(sizes, wname, level, 0, [0] * 4, params={"z_chunks": zc})
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), wname, level)) <= 1e-5
    assert orc.rel_l2(xr, x) <= 1e-5
    assert orc.rel_l2(xc, orc.rec_direct(c.astype(np.complex128), wname, False)) <= 1e-5


def test_emulated_ranks_generic_kernels_and_repeat_calls():
    """kernel_mode=1 (generic kernels, gather-form synthesis) and plan reuse across several pairs."""
    sizes, wname, level = (32, 24, 16, 12), "db3", 2
    x, y, xr, c, xc, tr = _run(sizes, wname, level, 0, [0, 0, 0], kernel_mode=1)
    assert not tr.plan.separable
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), wname, level)) <= 1e-5
    assert orc.rel_l2(xr, x) <= 1e-5
    assert orc.rel_l2(xc, orc.rec_direct(c.astype(np.complex128), wname, False)) <= 1e-5
    import torch
    xs = tr.scatter_input(nd.to_device(x).permute(3, 2, 1, 0).contiguous())
    for _ in range(3):                          # flags / events / buffers are reused consistently
        ys = tr.dec(xs, level)
        xs2 = tr.rec(ys)
    got = nd.to_host(torch.cat(xs2, dim=0).permute(3, 2, 1, 0))
    assert orc.rel_l2(got, x) <= 1e-5


def test_emulated_ranks_dilated_atrous_halos():
    """a-trous mode across ranks: halos of (L-1)*dil planes per level (level 3: 12 planes for 5-plane slabs)."""
    sizes, wname, level, dil = (24, 20, 20), "db2", 3, [1, 2, 4]
    x, y, xr, c, xc, tr = _run(sizes, wname, level, 1, [0] * 4, dil=dil)
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), wname, level, True, dilations=dil)) <= 1e-5
    assert orc.rel_l2(xr, x) <= 1e-5


def test_cfg4_row_length_schedule():
    """Rows of 256 elements (cfg4's row length: the 512-thread full-row synthesis instantiation) in the
    slab schedule, 4 planes per rank like cfg4 on 8 GPUs."""
    sizes, wname, level = (256, 24, 10, 16), "db4", 2
    x, y, xr, c, xc, tr = _run(sizes, wname, level, 0, [0] * 4, params={"rows_min_ctas": 0})
    assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), wname, level)) <= 1e-5
    assert orc.rel_l2(xr, x) <= 1e-5
    assert orc.rel_l2(xc, orc.rec_direct(c.astype(np.complex128), wname, False)) <= 1e-5


@pytest.mark.parametrize("sizes,wn,level,dtype", [((32, 24, 16, 12), "db4", 2, "complex64"), ((40, 20, 18), "db3", 2, "float64"),
                                                   ((64, 30), ["db2", "db4"], 2, "complex128")])
def test_host_arrays_through_the_multi_gpu_plan(sizes, wn, level, dtype):
    """'ngpus' option of the object API (what the MEX gateway's 'ngpus' plans call): whole host arrays in and out,
    every rank copies its own contiguous slab of x and of every band (nddwt_mplan_dec_host / rec_host).
    Ranks are emulated on one GPU through the 'devices' hook when the box has fewer GPUs."""
    n = min(_ngpus(), 3)
    devices = list(range(n)) if n >= 2 else [0, 0, 0]
    d = len(sizes)
    prec = "single" if dtype in ("complex64", "float32") else "double"
    tol = 1e-5 if prec == "single" else 1e-12
    cls = {2: nd.nd_dwt_2D, 3: nd.nd_dwt_3D, 4: nd.nd_dwt_4D}[d]
    o = cls(wn, list(sizes), "precision", prec, "compute", "mex", "devices", devices)
    x = orc.synth(sizes, dtype, 4)
    y = o.dec(x, level)
    wide = np.complex128 if np.iscomplexobj(x) else np.float64
    assert orc.rel_l2(y, orc.dec_direct(x.astype(wide), wn, level)) <= tol
    assert orc.rel_l2(o.rec(y), x) <= tol


def _ngpus():
    import torch
    return torch.cuda.device_count()


def test_real_devices_one_process():
    n = min(_ngpus(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the emulated-rank tests cover the schedule on one)")
    for sizes, wname, level, l2 in [((32, 24, 16, 4 * n), "db4", 3, 0), ((40, 20, 12, 2 * n + 1), "db2", 2, 1),
                                    ((64, 32, max(3 * n, 9)), "db4", 2, 0)]:
        x, y, xr, c, xc, tr = _run(sizes, wname, level, l2, list(range(n)))
        assert orc.rel_l2(y, orc.dec_direct(x.astype(np.complex128), wname, level, bool(l2))) <= 1e-5
        assert orc.rel_l2(xr, x) <= 1e-5
        assert orc.rel_l2(xc, orc.rec_direct(c.astype(np.complex128), wname, bool(l2))) <= 1e-5


def test_one_process_per_gpu_torchrun():
    """The bench's N>1 form: torchrun, one rank per GPU, CUDA-IPC inboxes + flags in peer memory (and the
    NCCL send/recv schedule of slab.SlabTransform as the second transport), every rank vs the oracle."""
    n = min(_ngpus(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    assert "multi-GPU parity OK" in p.stdout
