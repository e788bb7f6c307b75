"""Multi-GPU path: slabs along the LAST dimension, one process per GPU, periodic halo exchange.

The reference has no multi-device code; this module is the B200-native extension
BASELINE.json's north_star asks for.  Rank r owns a contiguous block of planes of the last
dimension of x and of every subband.  Per level:

  analysis   needs (L/2 - 1) planes of the approximation band from below and L/2 from above
             (L = taps along the last dim) -> one grouped send/recv on the periodic ring, then the
             slab kernel (nddwt_dec_level_slab);
  synthesis  is split at the slab dimension: stage 1 synthesises dims 1..d-1 locally and sums the
             band bits into two arrays u_lo / u_hi, their L/2 (below) and L/2 - 1 (above) halo planes
             are exchanged, stage 2 finishes along the last dim (nddwt_rec_level_slab_stage1/2).

When the halo is wider than a neighbour's slab (cfg4 at 8 GPUs: 4 planes per rank, 7 halo planes)
planes come from several ranks; `halo_sources` resolves every needed plane to (owner, local index).

`torch.distributed` is only the plumbing (NCCL on GPUs; gloo in the CPU tests); the compute engine
is the CUDA library.  The engine is pluggable so the host logic can be tested on CPU with a
numpy stand-in (tests/test_slab_gloo.py) -- the product default has no CPU engine.
"""
from __future__ import annotations

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def slab_partition(n_last, world):
    """Contiguous, at most +-1 ragged split of n_last planes over `world` ranks -> [(start, count)]."""
    base, rem = divmod(int(n_last), int(world))
    out, s = [], 0
    for r in range(world):
        c = base + (1 if r < rem else 0)
        out.append((s, c))
        s += c
    return out


def owner_of(plane, parts):
    for r, (s, c) in enumerate(parts):
        if s <= plane < s + c:
            return r, plane - s
    raise ValueError("plane out of range")


def halo_sources(n_last, world, rank, below, above):
    """Global planes a rank needs around its slab, resolved to (owner rank, local index).
    Returns (lo, hi): lists ordered by ascending global position (periodic)."""
    parts = slab_partition(n_last, world)
    s, c = parts[rank]
    lo = [owner_of((s - below + i) % n_last, parts) for i in range(below)]
    hi = [owner_of((s + c + i) % n_last, parts) for i in range(above)]
    return lo, hi


def _runs(entries):
    """Group consecutive (owner, idx) entries into (owner, first_idx, count, dst_offset) runs."""
    runs, i = [], 0
    while i < len(entries):
        o, idx = entries[i]
        j = i + 1
        while j < len(entries) and entries[j][0] == o and entries[j][1] == entries[j - 1][1] + 1:
            j += 1
        runs.append((o, idx, j - i, i))
        i = j
    return runs


class HaloExchanger:
    """Precomputed send/recv schedule for one (below, above) halo shape on the periodic ring."""

    def __init__(self, n_last, world, rank, below, above, group=None):
        self.n_last, self.world, self.rank, self.below, self.above, self.group = n_last, world, rank, below, above, group
        self.recv_lo = _runs(halo_sources(n_last, world, rank, below, above)[0])
        self.recv_hi = _runs(halo_sources(n_last, world, rank, below, above)[1])
        # what every other rank needs from me
        self.send = []   # (peer, first_idx, count, tag)
        for peer in range(world):
            plo, phi = halo_sources(n_last, world, peer, below, above)
            for which, runs in ((0, _runs(plo)), (1, _runs(phi))):
                for k, (o, idx, cnt, _) in enumerate(runs):
                    if o == rank:
                        self.send.append((peer, idx, cnt, which, k))
        # planes other ranks need from this one (= planes received by reverse_exchange)
        self.stage_planes = sum(cnt for peer, _, cnt, _, _ in self.send if peer != rank)

    def exchange(self, local, halo_lo, halo_hi):
        """local: [n_local, ...] contiguous; halo_lo: [below, ...]; halo_hi: [above, ...] (filled in place).
        Planes a rank needs from itself are copied locally."""
        ops, pending_self = [], []
        for which, runs, dst in ((0, self.recv_lo, halo_lo), (1, self.recv_hi, halo_hi)):
            for k, (o, idx, cnt, off) in enumerate(runs):
                if o == self.rank:
                    pending_self.append((dst, off, idx, cnt))
                else:
                    ops.append(dist.P2POp(dist.irecv, dst[off:off + cnt], o, group=self.group,
                                          tag=which * 64 + k))
        for peer, idx, cnt, which, k in self.send:
            if peer != self.rank:
                ops.append(dist.P2POp(dist.isend, local[idx:idx + cnt], peer, group=self.group, tag=which * 64 + k))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        for dst, off, idx, cnt in pending_self:
            dst[off:off + cnt].copy_(local[idx:idx + cnt])
        for r in reqs:
            r.wait()


def reverse_exchange(xch, over_lo, over_hi, stage):
    """Adjoint of HaloExchanger.exchange: over_lo / over_hi hold partial sums this rank computed for
    planes owned by other ranks (same plane sets as the halos of `xch`); they are sent to the owners.
    What other ranks computed for MY planes arrives in `stage` ([total, ...]).  Returns the list of
    (local_first_plane, count, stage_offset) runs the caller has to add into its result."""
    ops, adds, self_adds = [], [], []
    off = 0
    for peer, idx, cnt, which, k in xch.send:            # forward mode: I send planes idx.. to `peer`
        if peer == xch.rank:
            continue
        ops.append(dist.P2POp(dist.irecv, stage[off:off + cnt], peer, group=xch.group, tag=which * 64 + k))
        adds.append((idx, cnt, off))
        off += cnt
    for which, runs, src in ((0, xch.recv_lo, over_lo), (1, xch.recv_hi, over_hi)):
        for k, (o, idx, cnt, o_off) in enumerate(runs):  # forward mode: I receive these planes from `o`
            if o == xch.rank:
                self_adds.append((idx, cnt, src, o_off))
            else:
                ops.append(dist.P2POp(dist.isend, src[o_off:o_off + cnt], o, group=xch.group, tag=which * 64 + k))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    for r in reqs:
        r.wait()
    return adds, self_adds


class CudaSlabEngine:
    """Compute engine backed by libnddwt_b200.so (slab entry points of include/nddwt_b200.h)."""

    def __init__(self, local_sizes, global_last, wnames, dtype_code, pres_l2_norm, device_index):
        from ._lib import Plan
        self.plan = Plan(local_sizes, wnames, dtype_code, pres_l2_norm, device_index, global_last=global_last)
        self.device_index = device_index
        self.separable = self.plan.separable    # levels can be issued in parts (overlap with the exchange)

    def halo_planes(self, level):
        return self.plan.halo_planes(level)

    def _stream(self):
        return torch.cuda.current_stream(self.device_index).cuda_stream

    def dec_level(self, level, a_in, halo_lo, halo_hi, out_bands):
        self.plan.dec_level_slab(level, a_in.data_ptr(), halo_lo.data_ptr() if halo_lo is not None else None,
                                 halo_hi.data_ptr() if halo_hi is not None else None,
                                 [t.data_ptr() for t in out_bands], self._stream())

    def rec_stage1(self, level, in_bands, u_lo, u_hi):
        self.plan.rec_level_slab_stage1(level, [t.data_ptr() for t in in_bands], u_lo.data_ptr(), u_hi.data_ptr(),
                                        self._stream())

    # part-wise entry points (overlap of the halo exchange with compute, see SlabTransform._dec_overlapped)
    def dec_level_part(self, level, part, a_in, halo_lo, halo_hi, out_bands):
        self.plan.dec_level_slab_part(level, part, a_in.data_ptr(), halo_lo.data_ptr(), halo_hi.data_ptr(),
                                      [t.data_ptr() for t in out_bands], self._stream())

    def rec_stage2_scatter(self, level, u_lo, u_hi, a_out, over_lo, over_hi):
        self.plan.rec_level_slab_stage2_scatter(level, u_lo.data_ptr(), u_hi.data_ptr(), a_out.data_ptr(),
                                                over_lo.data_ptr(), over_hi.data_ptr(), self._stream())

    def accumulate(self, dst, src):
        self.plan.accumulate(dst.data_ptr(), src.data_ptr(), dst.numel(), self._stream())

    def rec_stage1_part(self, level, part, in_bands, u_lo, u_hi):
        self.plan.rec_level_slab_stage1_part(level, part, [t.data_ptr() for t in in_bands], u_lo.data_ptr(),
                                             u_hi.data_ptr(), self._stream())

    def rec_stage2(self, level, u_lo, u_hi, halo_lo, halo_hi, a_out):
        self.plan.rec_level_slab_stage2(level, u_lo.data_ptr(), u_hi.data_ptr(),
                                        halo_lo.data_ptr() if halo_lo is not None else None,
                                        halo_hi.data_ptr() if halo_hi is not None else None, a_out.data_ptr(),
                                        self._stream())


class SlabTransform:
    """dec / rec of one rank's slab.  Tensors are [n_local, N_{d-1}, ..., N_1] C-contiguous (i.e. the
    column-major MATLAB array restricted to the rank's planes of the last dim); the coefficient slab
    is [nb, n_local, ...] with the reference's band order (deepest level first)."""

    def __init__(self, sizes, wnames, level_max, engine, taps_last, rank, world, group=None, device="cpu",
                 dtype=None, overlap=True, scatter=True):
        self.sizes = tuple(int(s) for s in sizes)
        self.d = len(self.sizes)
        self.rank, self.world, self.group = rank, world, group
        self.parts = slab_partition(self.sizes[-1], world)
        self.start, self.n_local = self.parts[rank]
        self.engine = engine
        self.L = int(taps_last)
        self.device, self.dtype = device, dtype
        self.local_shape = (self.n_local,) + tuple(reversed(self.sizes[:-1]))
        lo_d, hi_d = self.L // 2 - 1, self.L // 2
        self.x_dec = HaloExchanger(self.sizes[-1], world, rank, lo_d, hi_d, group)
        self.x_rec = HaloExchanger(self.sizes[-1], world, rank, self.L // 2, self.L // 2 - 1, group)
        plane = self.local_shape[1:]
        mk = lambda n, mult=1: torch.empty((max(n * mult, 1),) + plane, dtype=dtype, device=device)
        self.h_dec = (mk(lo_d), mk(hi_d))
        self.h_rec = (mk(self.L // 2, 2), mk(self.L // 2 - 1, 2))     # u_lo planes then u_hi planes
        self.approx = [torch.empty(self.local_shape, dtype=dtype, device=device) for _ in range(2)]
        self.u = [torch.empty(self.local_shape, dtype=dtype, device=device) for _ in range(2)]
        # overlap of exchange and compute: needs the part-wise engine entry points, CUDA tensors and >1 rank
        self.overlap = (world > 1 and getattr(engine, "separable", False) and torch.device(device).type == "cuda"
                        and bool(overlap))
        if self.overlap:
            self.comm_stream = torch.cuda.Stream(device=device, priority=-1)
            self.u_hi2 = [self.u[1], torch.empty(self.local_shape, dtype=dtype, device=device)]
            self.h_rec2 = [self.h_rec, (mk(self.L // 2, 2), mk(self.L // 2 - 1, 2))]
            # scatter-form synthesis exchange: overhang partial sums (same plane sets as the analysis halos)
            self.scatter = bool(scatter)
            self.over = (mk(lo_d), mk(hi_d))
            self.stage = mk(self.x_dec.stage_planes)

    def num_bands(self, level):
        nd = 1 << self.d
        return nd + (nd - 1) * (level - 1)

    # ---- overlapped schedules (multi-GPU): the exchange of the next level's halos runs on a side
    # stream while the tile pass for the bands that do not feed it is still computing ----------
    def _dec_overlapped(self, x_local, level, out):
        nd = 1 << self.d
        cur = torch.cuda.current_stream(x_local.device)
        comm = self.comm_stream
        hl, hh = self.h_dec
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            self.x_dec.exchange(x_local, hl, hh)
        a_in = x_local
        for j in range(1, level + 1):
            start = (nd - 1) * (level - j)
            bands = [None] + [out[start + b] for b in range(1, nd)]
            bands[0] = out[0] if j == level else self.approx[j & 1]
            cur.wait_stream(comm)                                   # halos of a_in have arrived
            self.engine.dec_level_part(j, 1, a_in, hl, hh, bands)   # last-dim pass (reads the halos)
            self.engine.dec_level_part(j, 2, a_in, hl, hh, bands)   # bands 0..nd/2-1, incl. the approximation
            if j < level:
                comm.wait_stream(cur)
                with torch.cuda.stream(comm):
                    self.x_dec.exchange(bands[0], hl, hh)           # next level's halos, overlapped with ...
            self.engine.dec_level_part(j, 3, a_in, hl, hh, bands)   # ... the remaining detail bands
            a_in = bands[0]
        return out

    def _rec_overlapped_scatter(self, coeffs, out):
        """Synthesis with the scatter-form exchange: per level one array of overhang partial sums moves
        (L-1 planes) instead of the halos of u_lo and u_hi; the detail-only half of the NEXT level's
        stage 1 runs while they are in flight."""
        nd = 1 << self.d
        nb = coeffs.shape[0]
        level = 1 + (nb - nd) // (nd - 1)
        cur = torch.cuda.current_stream(coeffs.device)
        comm = self.comm_stream
        u_lo = self.u[0]
        over_lo, over_hi = self.over

        def bands_of(j, a):
            start = (nd - 1) * (level - j)
            return [a] + [coeffs[start + b] for b in range(1, nd)]

        a = coeffs[0]
        self.engine.rec_stage1_part(level, 2, bands_of(level, a), u_lo, self.u_hi2[level & 1])
        for j in range(level, 0, -1):
            k = j & 1
            dst = out if j == 1 else self.approx[j & 1]
            self.engine.rec_stage1_part(j, 1, bands_of(j, a), u_lo, self.u_hi2[k])
            self.engine.rec_stage2_scatter(j, u_lo, self.u_hi2[k], dst, over_lo, over_hi)
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                adds, self_adds = reverse_exchange(self.x_dec, over_lo, over_hi, self.stage)
            if j > 1:   # detail half of the next level: independent of dst, overlaps with the exchange
                self.engine.rec_stage1_part(j - 1, 2, bands_of(j - 1, coeffs[0]), u_lo, self.u_hi2[(j - 1) & 1])
            cur.wait_stream(comm)
            for idx, cnt, off in adds:
                self.engine.accumulate(dst[idx:idx + cnt], self.stage[off:off + cnt])
            for idx, cnt, src, o_off in self_adds:
                self.engine.accumulate(dst[idx:idx + cnt], src[o_off:o_off + cnt])
            a = dst
        return out

    def _rec_overlapped(self, coeffs, out):
        if getattr(self, "scatter", False):
            return self._rec_overlapped_scatter(coeffs, out)
        nd = 1 << self.d
        nb = coeffs.shape[0]
        level = 1 + (nb - nd) // (nd - 1)
        cur = torch.cuda.current_stream(coeffs.device)
        comm = self.comm_stream
        below, above = self.L // 2, self.L // 2 - 1
        u_lo = self.u[0]

        def bands_of(j, a):
            start = (nd - 1) * (level - j)
            return [a] + [coeffs[start + b] for b in range(1, nd)]

        def stage1b(j):      # detail-only half: independent of the previous level's result
            k = j & 1
            self.engine.rec_stage1_part(j, 2, bands_of(j, coeffs[0]), u_lo, self.u_hi2[k])
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                hl, hh = self.h_rec2[k]
                self.x_rec.exchange(self.u_hi2[k], hl[below:2 * below], hh[above:2 * above])

        a = coeffs[0]
        stage1b(level)
        for j in range(level, 0, -1):
            k = j & 1
            hl, hh = self.h_rec2[k]
            dst = out if j == 1 else self.approx[j & 1]
            self.engine.rec_stage1_part(j, 1, bands_of(j, a), u_lo, self.u_hi2[k])
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                self.x_rec.exchange(u_lo, hl[:below], hh[:above])
                ev = torch.cuda.Event()
                ev.record(comm)
            if j > 1:
                stage1b(j - 1)                                       # overlaps with the u_lo exchange of level j
            cur.wait_event(ev)
            self.engine.rec_stage2(j, u_lo, self.u_hi2[k], hl, hh, dst)
            a = dst
        return out

    def dec(self, x_local, level, out=None):
        nd = 1 << self.d
        nb = self.num_bands(level)
        if out is None:
            out = torch.empty((nb,) + self.local_shape, dtype=x_local.dtype, device=x_local.device)
        if self.overlap:
            return self._dec_overlapped(x_local, level, out)
        a_in = x_local
        for j in range(1, level + 1):
            start = (nd - 1) * (level - j)           # slot arithmetic of mex/nddwt.c:209-210,226
            bands = [None] + [out[start + b] for b in range(1, nd)]
            bands[0] = out[0] if j == level else self.approx[j & 1]
            if self.world > 1:
                self.x_dec.exchange(a_in, self.h_dec[0], self.h_dec[1])
                self.engine.dec_level(j, a_in, self.h_dec[0], self.h_dec[1], bands)
            else:
                self.engine.dec_level(j, a_in, None, None, bands)
            a_in = bands[0]
        return out

    def rec(self, coeffs, out=None):
        nd = 1 << self.d
        nb = coeffs.shape[0]
        level = 1 + (nb - nd) // (nd - 1)
        if out is None:
            out = torch.empty(self.local_shape, dtype=coeffs.dtype, device=coeffs.device)
        if self.overlap:
            return self._rec_overlapped(coeffs, out)
        a = coeffs[0]
        below, above = self.L // 2, self.L // 2 - 1
        for j in range(level, 0, -1):
            start = (nd - 1) * (level - j)
            bands = [a] + [coeffs[start + b] for b in range(1, nd)]
            dst = out if j == 1 else self.approx[j & 1]
            self.engine.rec_stage1(j, bands, self.u[0], self.u[1])
            if self.world > 1:
                hl, hh = self.h_rec
                # u_lo halo planes first, then u_hi (layout of nddwt_rec_level_slab_stage2)
                self.x_rec.exchange(self.u[0], hl[:below], hh[:above])
                self.x_rec.exchange(self.u[1], hl[below:2 * below], hh[above:2 * above])
                self.engine.rec_stage2(j, self.u[0], self.u[1], hl, hh, dst)
            else:
                self.engine.rec_stage2(j, self.u[0], self.u[1], None, None, dst)
            a = dst
        return out


# ---------------------------------------------------------------------------------------------
# Peer-memory transport: the exchange lives inside the library (nddwt_mplan_*, csrc/nddwt_multi.cu):
# halo planes are pushed straight into the neighbours' inboxes with copy-engine peer copies, ranks
# order themselves with flags in peer memory (one process per GPU) or CUDA events (one process for
# all GPUs).  torch.distributed is only used once, to all-gather the IPC handles.
class PeerSlabTransform:
    """One rank of a one-process-per-GPU job; same tensor layout as SlabTransform
    ([n_local, N_{d-1}, ..., N_1] slabs, coefficient slab [nb, n_local, ...])."""

    transport = "peer-memory pushes (CUDA IPC + copy engines), flags in peer memory"

    def __init__(self, sizes, wnames, dtype_code, pres_l2_norm, rank, world, device_index, group=None, dilations=None):
        from ._lib import MultiPlan, lib
        import ctypes
        self.sizes = tuple(int(s) for s in sizes)
        self.d = len(self.sizes)
        self.rank, self.world, self.device_index = rank, world, device_index
        # Every rank takes part in every collective below whatever happens locally, and all ranks fail together:
        # a rank that cannot create or map its plan must not leave the others waiting in an all-gather.
        err, blob = None, b""
        try:
            self.plan = MultiPlan(self.sizes, wnames, dtype_code, pres_l2_norm, rank=rank, world=world, device=device_index)
            if dilations is not None:
                self.plan.set_dilations(dilations)
            n = int(lib().nddwt_mplan_export_size())
            buf = (ctypes.c_char * n)()
            from ._lib import check
            check(lib().nddwt_mplan_export(self.plan.handle, buf))
            blob = bytes(buf)
        except Exception as exc:   # noqa: BLE001
            err = "rank %d: %s" % (rank, str(exc)[:200])
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, (err, blob), group=group)
            errs = [e for e, _ in gathered if e]
            if not errs:
                try:
                    joined = b"".join(b for _, b in gathered)
                    cbuf = (ctypes.c_char * len(joined)).from_buffer_copy(joined)
                    check(lib().nddwt_mplan_import(self.plan.handle, cbuf))
                except Exception as exc:   # noqa: BLE001
                    err = "rank %d: %s" % (rank, str(exc)[:200])
                gathered = [None] * world
                dist.all_gather_object(gathered, err, group=group)
                errs = [e for e in gathered if e]
            if errs:
                raise RuntimeError("peer-memory plan could not be set up: " + "; ".join(errs))
        elif err:
            raise RuntimeError(err)
        self.start, self.n_local = self.plan.slab(rank)
        self.parts = [self.plan.slab(r) for r in range(world)]
        self.local_shape = (self.n_local,) + tuple(reversed(self.sizes[:-1]))
        self.overlap = self.plan.separable
        self.scatter = self.plan.separable

    def num_bands(self, level):
        nd = 1 << self.d
        return nd + (nd - 1) * (level - 1)

    def _stream(self):
        return [torch.cuda.current_stream(self.device_index).cuda_stream]

    def dec(self, x_local, level, out=None):
        if out is None:
            out = torch.empty((self.num_bands(level),) + self.local_shape, dtype=x_local.dtype, device=x_local.device)
        assert x_local.is_contiguous() and out.is_contiguous() and tuple(x_local.shape) == self.local_shape
        self.plan.dec([x_local.data_ptr()], [out.data_ptr()], level, self._stream())
        return out

    def rec(self, coeffs, out=None):
        nd = 1 << self.d
        level = 1 + (coeffs.shape[0] - nd) // (nd - 1)
        if out is None:
            out = torch.empty(self.local_shape, dtype=coeffs.dtype, device=coeffs.device)
        assert coeffs.is_contiguous() and out.is_contiguous()
        self.plan.rec([coeffs.data_ptr()], [out.data_ptr()], level, self._stream())
        return out


class MultiGpuTransform:
    """One process drives several GPUs (or several emulated ranks on one GPU: repeat a device index).
    x / coefficient slabs are lists with one tensor per rank on that rank's device."""

    def __init__(self, sizes, wnames, dtype_code, pres_l2_norm, devices, dilations=None, kernel_mode=0):
        from ._lib import MultiPlan
        self.sizes = tuple(int(s) for s in sizes)
        self.d = len(self.sizes)
        self.devices = list(devices)
        self.world = len(self.devices)
        self.plan = MultiPlan(self.sizes, wnames, dtype_code, pres_l2_norm, devices=self.devices)
        if kernel_mode:
            self.plan.set_kernel_mode(kernel_mode)
        if dilations is not None:
            self.plan.set_dilations(dilations)
        self.parts = [self.plan.slab(r) for r in range(self.world)]
        self.local_shapes = [(c,) + tuple(reversed(self.sizes[:-1])) for (_, c) in self.parts]

    def num_bands(self, level):
        nd = 1 << self.d
        return nd + (nd - 1) * (level - 1)

    def scatter_input(self, x_full_base):
        """[N_d, ..., N_1] tensor (any device) -> list of per-rank slabs on the ranks' devices."""
        return [x_full_base[s:s + c].to("cuda:%d" % dev).contiguous() for (s, c), dev in zip(self.parts, self.devices)]

    def dec(self, xs, level, outs=None):
        nb = self.num_bands(level)
        if outs is None:
            outs = [torch.empty((nb,) + shp, dtype=x.dtype, device=x.device) for x, shp in zip(xs, self.local_shapes)]
        for dev in set(self.devices):
            torch.cuda.synchronize(dev)          # plan-owned streams: inputs must be complete
        self.plan.dec([x.data_ptr() for x in xs], [o.data_ptr() for o in outs], level)
        self.plan.sync()
        return outs

    def rec(self, coeffs, outs=None):
        nd = 1 << self.d
        level = 1 + (coeffs[0].shape[0] - nd) // (nd - 1)
        if outs is None:
            outs = [torch.empty(shp, dtype=c.dtype, device=c.device) for c, shp in zip(coeffs, self.local_shapes)]
        for dev in set(self.devices):
            torch.cuda.synchronize(dev)
        self.plan.rec([c.data_ptr() for c in coeffs], [o.data_ptr() for o in outs], level)
        self.plan.sync()
        return outs
