"""CPU tests of the multi-GPU host logic (slab partition, multi-hop periodic halo exchange, level
loop, split synthesis) with torch.distributed/gloo, world_size 2 and 3.  The compute engine here
is a numpy stand-in built on the oracle (test infrastructure); on the GPU box the same
SlabTransform drives the CUDA slab kernels (tests/test_gpu_slab.py)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nddwt_oracle as orc

slab = importlib.import_module("non-decimated_wavelets_b200.slab")


def test_partition_and_halo_sources():
    assert slab.slab_partition(32, 8) == [(4 * r, 4) for r in range(8)]
    assert slab.slab_partition(7, 3) == [(0, 3), (3, 2), (5, 2)]
    # cfg4 at 8 GPUs: 4 planes per rank, db4 analysis halo 3 below / 4 above -> reaches 2 ranks up
    lo, hi = slab.halo_sources(32, 8, 0, 3, 4)
    assert lo == [(7, 1), (7, 2), (7, 3)] and hi == [(1, 0), (1, 1), (1, 2), (1, 3)]
    lo, hi = slab.halo_sources(32, 8, 0, 4, 3)     # synthesis
    assert lo == [(7, 0), (7, 1), (7, 2), (7, 3)] and hi == [(1, 0), (1, 1), (1, 2)]
    lo, hi = slab.halo_sources(7, 3, 1, 3, 4)      # halo wider than the slabs: multi-hop + self
    assert lo == [(0, 0), (0, 1), (0, 2)] and hi == [(2, 0), (2, 1), (0, 0), (0, 1)]


def test_library_routing_matches_python_routing():
    """The C library's halo routing (nddwt_slab_route, used by the peer-memory schedule of nddwt_mplan_*)
    resolves every halo plane to the same (owner, local index) as slab.halo_sources."""
    _lib = importlib.import_module("non-decimated_wavelets_b200._lib")
    for n, world in [(32, 8), (7, 3), (48, 8), (9, 2), (5, 5), (16, 1)]:
        for below, above in [(3, 4), (4, 3), (0, 1), (9, 10), (12, 16)]:
            for rank in range(world):
                lo, hi = slab.halo_sources(n, world, rank, below, above)
                for which, ref in ((0, lo), (1, hi)):
                    got = []
                    for owner, idx, cnt, off in _lib.slab_route(n, world, rank, which, below, above):
                        assert off == len(got)
                        got += [(owner, idx + i) for i in range(cnt)]
                    assert got == ref, (n, world, rank, which, below, above)


class NumpyEngine:
    """Slab compute stand-in: periodic filtering along dims 1..d-1, halo-fed along the last dim."""

    def __init__(self, wnames, pres_l2):
        self.filt = [orc.wave_filters(w) for w in wnames]
        self.d = len(wnames)
        self.sd = 1 / np.sqrt(2) if pres_l2 else 1.0
        self.sr = 1 / np.sqrt(2) if pres_l2 else 0.5

    @staticmethod
    def _np(t):      # [n_local, N_{d-1}..N_1] tensor -> MATLAB-shaped numpy (N_1..N_{d-1}, n_local)
        return t.numpy().transpose(*reversed(range(t.dim())))

    def dec_level(self, level, a_in, halo_lo, halo_hi, out_bands):
        lo, hi = self.filt[-1]
        L = len(lo)
        a = self._np(a_in)
        ext = np.concatenate([self._np(halo_lo)[..., :L // 2 - 1], a, self._np(halo_hi)[..., :L // 2]], axis=-1)
        n = a.shape[-1]
        cur = []
        for g in (lo, hi):        # y[m] = sum_k g[k] ext[m + below - (k - L/2)]
            acc = np.zeros_like(a)
            for k in range(L):
                s0 = (L // 2 - 1) - (k - L // 2)
                acc += self.sd * g[k] * ext[..., s0:s0 + n]
            cur.append(acc)
        # remaining dims, periodic; band bit order: dim 1 is the LSB
        arrays = {0: cur[0], 1: cur[1]}      # key = bits for dims already processed (last dim only)
        for i in reversed(range(self.d - 1)):
            lo_i, hi_i = self.filt[i]
            nxt = {}
            for key, arr in arrays.items():
                nxt[key * 2 + 0] = self.sd * orc._filt_axis(arr, lo_i, i, +1)
                nxt[key * 2 + 1] = self.sd * orc._filt_axis(arr, hi_i, i, +1)
            arrays = nxt
        for b, t in enumerate(out_bands):
            self._np(t)[...] = arrays[b]

    def rec_stage1(self, level, in_bands, u_lo, u_hi):
        cur = [self._np(t).copy() for t in in_bands]
        for i in range(self.d - 1):                  # synthesise dims 1..d-1 (pairs differ in the LSB first)
            lo_i, hi_i = self.filt[i]
            cur = [self.sr * (orc._filt_axis(cur[2 * j], lo_i, i, -1) + orc._filt_axis(cur[2 * j + 1], hi_i, i, -1))
                   for j in range(len(cur) // 2)]
        self._np(u_lo)[...] = cur[0]
        self._np(u_hi)[...] = cur[1]

    def rec_stage2(self, level, u_lo, u_hi, halo_lo, halo_hi, a_out):
        lo, hi = self.filt[-1]
        L = len(lo)
        below, above = L // 2, L // 2 - 1
        hl, hh = self._np(halo_lo), self._np(halo_hi)
        n = u_lo.shape[0]
        acc = np.zeros_like(self._np(u_lo))
        for g, u, off in ((lo, u_lo, 0), (hi, u_hi, 1)):
            ext = np.concatenate([hl[..., off * below:(off + 1) * below], self._np(u),
                                  hh[..., off * above:(off + 1) * above]], axis=-1)
            for k in range(L):   # x[m] = sum_k g[k] ext[m + below + (k - L/2)]
                s0 = below + (k - L // 2)
                acc += self.sr * g[k] * ext[..., s0:s0 + n]
        self._np(a_out)[...] = acc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sizes, wnames, level, l2, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = orc.synth(sizes, np.complex128, 42)                     # same seeded global array on every rank
        parts = slab.slab_partition(sizes[-1], world)
        s, c = parts[rank]
        xl = torch.from_numpy(np.ascontiguousarray(x[..., s:s + c].transpose(*reversed(range(len(sizes))))))
        eng = NumpyEngine(wnames, l2)
        L = len(orc.wave_filters(wnames[-1])[0])
        tr = slab.SlabTransform(sizes, wnames, level, eng, L, rank, world, device="cpu", dtype=torch.complex128)
        y = tr.dec(xl, level)
        xr = tr.rec(y)
        yo = orc.dec_direct(x, wnames, level, l2)                   # [sizes, nb]
        y_np = y.numpy().transpose(*reversed(range(y.dim())))       # (N1..n_local, nb)
        e_dec = orc.rel_l2(y_np, yo[..., s:s + c, :])
        e_rec = orc.rel_l2(xr.numpy().transpose(*reversed(range(xr.dim()))), x[..., s:s + c])
        q.put((rank, e_dec, e_rec))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,sizes,wnames,level,l2", [
    (2, (12, 10, 16), ["db2", "db1", "db4"], 2, False),
    (2, (8, 6, 6, 8), ["db2", "db2", "db2", "db2"], 2, True),
    (3, (10, 8, 7), ["db1", "db2", "db3"], 3, False),              # ragged slabs 3/2/2, halo wider than a slab
    (2, (16, 12), ["db3", "db1"], 2, True),                        # Haar along the slab dim: one-sided halos
])
def test_slab_transform_world(world, sizes, wnames, level, l2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sizes, wnames, level, l2, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = [q.get(timeout=5) for _ in range(world)]
    for rank, e_dec, e_rec in res:
        assert e_dec < 1e-12 and e_rec < 1e-12, (rank, e_dec, e_rec)


def _adj_worker(rank, world, port, n_last, below, above, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = slab.slab_partition(n_last, world)
        s, c = parts[rank]
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn((c, 5), generator=g, dtype=torch.float64)
        ylo = torch.randn((max(below, 1), 5), generator=g, dtype=torch.float64)
        yhi = torch.randn((max(above, 1), 5), generator=g, dtype=torch.float64)
        xch = slab.HaloExchanger(n_last, world, rank, below, above)
        hlo, hhi = torch.zeros_like(ylo), torch.zeros_like(yhi)
        xch.exchange(x, hlo, hhi)                                  # forward: halos of x
        lhs = (hlo[:below] * ylo[:below]).sum() + (hhi[:above] * yhi[:above]).sum()
        stage = torch.zeros((max(xch.stage_planes, 1), 5), dtype=torch.float64)
        adds, self_adds = slab.reverse_exchange(xch, ylo, yhi, stage)   # adjoint: scatter-add of y
        z = torch.zeros_like(x)
        for idx, cnt, off in adds:
            z[idx:idx + cnt] += stage[off:off + cnt]
        for idx, cnt, src, o_off in self_adds:
            z[idx:idx + cnt] += src[o_off:o_off + cnt]
        rhs = (x * z).sum()
        both = torch.stack([lhs, rhs])
        dist.all_reduce(both)
        q.put((rank, float(both[0]), float(both[1])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_last,below,above", [(2, 16, 3, 4), (3, 7, 3, 4), (2, 8, 0, 1)])
def test_reverse_exchange_is_adjoint_of_exchange(world, n_last, below, above):
    """<exchange(x), y> == <x, reverse_exchange(y)> summed over ranks (scatter-form synthesis exchange)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_adj_worker, args=(r, world, port, n_last, below, above, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for _ in range(world):
        rank, lhs, rhs = q.get(timeout=5)
        assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs)), (rank, lhs, rhs)
