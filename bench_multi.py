"""bench_multi.py -- the N>1 leg of bench.py: one process per GPU (torchrun), slabs along the last
dimension.  Strong scaling on the named workload (cfg4 by default); the same-workload curve against the
N=1 line (cfg5, sharded) is timed beside it and reported in `config`.

Transport of the halo planes (`--transport`):
  peer  (default) the library's multi-GPU plan (nddwt_mplan_*, csrc/nddwt_multi.cu): copy-engine pushes into
        peer memory mapped through CUDA IPC, flags in peer memory; NCCL only carries the IPC handles once;
  nccl  the round-1 schedule: torch.distributed batch_isend_irecv per level (slab.SlabTransform).
If the peer plan cannot be set up on a box the run falls back to nccl and says so in `config`.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


COMM_STREAMS = 0
Z_CHUNKS = 0


def _make_transform(slab, _lib, nd, transport, sizes, wname, level, dtype, rank, world, local_rank, dev):
    import torch
    d = len(sizes)
    wn = [wname] * d
    code = _lib.NDDWT_C64 if dtype == "complex64" else _lib.NDDWT_C128
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    if transport == "peer":
        tr = slab.PeerSlabTransform(sizes, wn, code, 0, rank, world, local_rank)
        if COMM_STREAMS:
            tr.plan.set_param("comm_streams", COMM_STREAMS)
        if Z_CHUNKS:
            tr.plan.set_param("z_chunks", Z_CHUNKS)
        return tr, tr.plan
    L = len(nd.wave_filters(wname)[0])
    c0 = slab.slab_partition(sizes[-1], world)[rank][1]
    eng = slab.CudaSlabEngine(tuple(sizes[:-1]) + (c0,), sizes[-1], wn, code, 0, local_rank)
    tr = slab.SlabTransform(sizes, wn, level, eng, L, rank, world, device=dev, dtype=tdt)
    return tr, eng.plan


def _parity(slab, _lib, nd, transport, wl, rank, world, local_rank, dev):
    """Oracle comparison of this transport's schedule on a reduced array with the workload's slab geometry
    (same last dimension, same wavelet and levels): rank 0 checks its slab of dec against the oracle's
    full-array transform; every rank contributes to the reconstruction error."""
    import torch
    import torch.distributed as dist
    from oracle import nddwt_oracle as orc
    sizes, wname, level, dtype = wl
    small = tuple(min(s, 64) for s in sizes[:2]) + tuple(min(s, 8) for s in sizes[2:-1]) + (sizes[-1],)
    d = len(small)
    x = orc.synth(small, np.complex64 if dtype == "complex64" else np.complex128, 7)
    tr, _ = _make_transform(slab, _lib, nd, transport, small, wname, level, dtype, rank, world, local_rank, dev)
    s, c = tr.parts[rank]
    xl = torch.from_numpy(np.ascontiguousarray(x[..., s:s + c].transpose(*reversed(range(d))))).to(dev)
    y = tr.dec(xl, level)
    xr = tr.rec(y)
    torch.cuda.synchronize()
    e_rec = orc.rel_l2(xr.cpu().numpy().transpose(*reversed(range(d))), x[..., s:s + c])
    e_dec = 0.0
    if rank == 0:
        yo = orc.dec_direct(x.astype(np.complex128), [wname] * d, level, False)
        e_dec = orc.rel_l2(y.cpu().numpy().transpose(*reversed(range(d + 1))), yo[..., s:s + c, :])
    t = torch.tensor([e_dec, e_rec], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del tr
    return float(t[0]), float(t[1]), list(small)


def _time_pairs(tr, x, y, xr, level, steps, warmup, dev):
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        tr.dec(x, level, out=y)
        tr.rec(y, out=xr)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.dec(x, level, out=y)
        tr.rec(y, out=xr)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(ms[0]) / steps


def _time_e2e(tr, x, y, xr, level, steps, dev):
    """The same pair with HOST slabs: every step copies the rank's x slab in, transforms, copies the
    coefficient slab out, copies it back in, reconstructs and copies x out (the N=1 e2e of bench.py, sharded)."""
    import torch
    import torch.distributed as dist
    hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    hy = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
    hx2 = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    hx.copy_(x)

    def step():
        x.copy_(hx, non_blocking=True)
        tr.dec(x, level, out=y)
        hy.copy_(y, non_blocking=True)
        y.copy_(hy, non_blocking=True)
        tr.rec(y, out=xr)
        hx2.copy_(xr, non_blocking=True)

    step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    err = torch.linalg.vector_norm(hx2.to(dev) - hx.to(dev)) ** 2
    nrm = torch.linalg.vector_norm(hx.to(dev)) ** 2
    both = torch.stack([err, nrm]).to(torch.float64)
    dist.all_reduce(both)
    dist.barrier()
    nbytes = (x.numel() + y.numel()) * x.element_size()
    return float(ms[0]) / steps, nbytes, float(torch.sqrt(both[0] / both[1]))


def _time_loop(tr, plan, x, y, xr, level, iters, dev, d):
    """x <- rec(shrink(dec(x))), `iters` times, thresholds = median |d_j| per level (agreed through rank 0)."""
    import torch
    import torch.distributed as dist
    nd_b = 1 << d
    tr.dec(x, level, out=y)
    meds = []
    for j in range(1, level + 1):
        band = y[(nd_b - 1) * (level - j) + nd_b - 1].reshape(-1)
        meds.append(band[:: max(1, band.numel() // (1 << 20))].abs().median())
    t = torch.stack(meds).to(torch.float64)
    dist.broadcast(t, 0)
    table = np.repeat(t.cpu().numpy()[:, None], nd_b, axis=1)
    plan.set_shrink(table)
    xw = x.clone()
    for _ in range(2):
        tr.dec(xw, level, out=y)
        tr.rec(y, out=xw)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        tr.dec(xw, level, out=y)
        tr.rec(y, out=xw)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    plan.set_shrink(None)
    zero = (y[1:].reshape(-1)[:: 997] == 0).float().mean().to(torch.float64)
    dist.all_reduce(zero)
    return float(ms[0]) / iters, float(zero) / dist.get_world_size()


def _alloc(tr, level, dtype, dev, rank):
    import torch
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    rdt = torch.float32 if dtype == "complex64" else torch.float64
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.view_as_complex(torch.randn(tr.local_shape + (2,), generator=g, device=dev, dtype=rdt))
    nb = tr.num_bands(level)
    y = torch.empty((nb,) + tr.local_shape, dtype=tdt, device=dev)
    xr = torch.empty_like(x)
    return x, y, xr, nb


def run_multi(args, wl_name, wl):
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, peaks, WORKLOADS
    slab = importlib.import_module("non-decimated_wavelets_b200.slab")
    _lib = importlib.import_module("non-decimated_wavelets_b200._lib")
    import nddwt_b200 as nd

    sizes, wname, level, dtype = wl[:4]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)   # nccl transport: its kernels must not queue behind the tile kernels
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, pg_options=opts)

    transport = getattr(args, "transport", "peer")
    global COMM_STREAMS, Z_CHUNKS
    COMM_STREAMS = int(getattr(args, "comm_streams", 0) or 0)
    Z_CHUNKS = int(getattr(args, "z_chunks", 0) or 0)
    fallback = None
    if transport == "peer":
        # every rank must agree on the transport: any failure anywhere switches all ranks to nccl
        ok = 1
        try:
            e_dec, e_rec, small = _parity(slab, _lib, nd, "peer", (sizes, wname, level, dtype), rank, world, local_rank, dev)
            if not (e_dec <= 1e-5 and e_rec <= 1e-5):
                raise RuntimeError("peer transport parity %.2e / %.2e" % (e_dec, e_rec))
        except Exception as exc:   # noqa: BLE001
            ok, fallback = 0, "%s: %s" % (type(exc).__name__, str(exc)[:160])
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t[0]) == 0:
            transport = "nccl"
            fallback = fallback or "another rank could not set up the peer plan"
    if transport == "nccl":
        e_dec, e_rec, small = _parity(slab, _lib, nd, "nccl", (sizes, wname, level, dtype), rank, world, local_rank, dev)

    L = len(nd.wave_filters(wname)[0])
    tr, plan = _make_transform(slab, _lib, nd, transport, sizes, wname, level, dtype, rank, world, local_rank, dev)
    x, y, xr, nb = _alloc(tr, level, dtype, dev, rank)
    warm = max(3, args.warmup)
    for _ in range(warm):
        tr.dec(x, level, out=y)
        tr.rec(y, out=xr)
    torch.cuda.synchronize()
    err = torch.linalg.vector_norm(xr - x) ** 2
    nrm = torch.linalg.vector_norm(x) ** 2
    both = torch.stack([err, nrm])
    dist.all_reduce(both)
    pr_err = float(torch.sqrt(both[0] / both[1]))

    split = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    dist.barrier()
    torch.cuda.synchronize()
    split[0].record()
    tr.dec(x, level, out=y)
    split[1].record()
    tr.rec(y, out=xr)
    split[2].record()
    torch.cuda.synchronize()
    dec_ms, rec_ms = split[0].elapsed_time(split[1]), split[1].elapsed_time(split[2])

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = plan.launches
    hb0 = plan.halo_bytes if transport == "peer" else 0
    ms_per_step = _time_pairs(tr, x, y, xr, level, args.steps, 0, dev)
    hb1 = plan.halo_bytes if transport == "peer" else 0
    clocks = sampler.stop() if sampler else None
    launches = plan.launches - l0
    # per-kind device times of this rank (events around every launch, second pass; not part of the timed value)
    kinds = ["analysis tile kernel", "synthesis tile kernel", "analysis last-dim pass", "synthesis last-dim pass + adds",
             "generic pass", "halo pushes (comm stream)"]
    ktimes = None
    try:
        for k in range(6):
            plan.kernel_time(k)
        plan.profile(True)
        psteps = max(2, min(args.steps, 5))
        for _ in range(psteps):
            tr.dec(x, level, out=y)
            tr.rec(y, out=xr)
        torch.cuda.synchronize()
        plan.profile(False)
        ktimes = {}
        for k in range(6):
            tot, cnt = plan.kernel_time(k)
            if cnt:
                ktimes[kinds[k]] = {"ms_per_step": tot / psteps, "launches_per_step": cnt / psteps}
        compute = sum(v["ms_per_step"] for n, v in ktimes.items() if not n.startswith("halo"))
        ktimes["compute_kernels_ms_per_step"] = compute
        ktimes["exposed_comm_ms_per_step"] = ms_per_step - compute
        dist.barrier()
    except Exception as exc:   # noqa: BLE001
        ktimes = {"error": str(exc)[:160]}
    nvox = int(np.prod(sizes))
    esize = np.dtype(dtype).itemsize
    value = nvox / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = peaks()
    pair_bytes = 2 * (1 + nb) * nvox * esize
    pair_gbs = pair_bytes / (ms_per_step * 1e-3) / 1e9
    plane_bytes = int(np.prod(sizes[:-1])) * esize
    if transport == "peer":
        halo_bytes = (hb1 - hb0) // max(args.steps, 1)
        timeouts = plan.wait_timeouts
    else:
        halo_bytes = level * (L - 1) * plane_bytes * (2 if getattr(tr, "scatter", False) else 3)
        timeouts = 0
    overlap, scatter, parts = bool(tr.overlap), bool(getattr(tr, "scatter", False)), list(tr.parts)
    del tr, plan, x, y, xr
    torch.cuda.empty_cache()

    # ---- the N=1 workload (cfg5) sharded the same way: same-workload scaling curve and end-to-end number
    same = None
    e2e = {"value": None, "unit": "Mvoxels/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
           "unavailable": "host slabs of the %s coefficient stack do not fit" % wl_name}
    sw_name = "cfg5"
    if not getattr(args, "no_same_workload", False):
        s_sizes, s_wname, s_level, s_dtype = WORKLOADS[sw_name][:4]
        try:
            if wl_name == sw_name:
                raise RuntimeError("headline workload is already %s" % sw_name)
            tr2, plan2 = _make_transform(slab, _lib, nd, transport, s_sizes, s_wname, s_level, s_dtype, rank, world,
                                         local_rank, dev)
            x2, y2, xr2, nb2 = _alloc(tr2, s_level, s_dtype, dev, rank)
            ms2 = _time_pairs(tr2, x2, y2, xr2, s_level, max(5, args.steps // 2), 3, dev)
            nv2 = int(np.prod(s_sizes))
            same = {"workload": sw_name, "sizes": list(s_sizes), "ms_per_step": ms2, "value": nv2 / (ms2 * 1e-3) / 1e6,
                    "unit": "Mvoxels/s", "planes_per_gpu": "/".join(str(c) for _, c in tr2.parts),
                    "pair_frac_per_gpu": 2 * (1 + nb2) * nv2 * esize / (ms2 * 1e-3) / 1e9 / world / peak}
            iters = int(getattr(args, "loop", 0) or 0)
            if iters > 0 and transport == "peer":
                # BASELINE configs[4]: 100 dec/rec pairs on 192x192x64x48, db4 and Haar, plan and buffers reused,
                # with the soft threshold of the iterative loop fused into the analysis stores
                lms, zf = _time_loop(tr2, plan2, x2, y2, xr2, s_level, iters, dev, len(s_sizes))
                same["iterative_loop_db4_J3"] = {"iters": iters, "ms_per_iter": lms, "Mvox_per_s": nv2 / lms / 1e3,
                                                 "zeroed_fraction_of_details": zf}
                h_sizes, h_wname, h_level, h_dtype = WORKLOADS["cfg5haar"][:4]
                tr3, plan3 = _make_transform(slab, _lib, nd, transport, h_sizes, h_wname, h_level, h_dtype, rank, world,
                                             local_rank, dev)
                y3 = y2[:tr3.num_bands(h_level)]
                hms = _time_pairs(tr3, x2, y3, xr2, h_level, max(10, iters // 2), 3, dev)
                lms, zf = _time_loop(tr3, plan3, x2, y3, xr2, h_level, iters, dev, len(h_sizes))
                nbh = tr3.num_bands(h_level)
                same["haar_level1"] = {"pair_ms": hms, "pair_Mvox_per_s": nv2 / hms / 1e3,
                                       "pair_frac_per_gpu": 2 * (1 + nbh) * nv2 * esize / (hms * 1e-3) / 1e9 / world / peak,
                                       "iterative_loop": {"iters": iters, "ms_per_iter": lms, "Mvox_per_s": nv2 / lms / 1e3,
                                                          "zeroed_fraction_of_details": zf}}
                del tr3, plan3
            if not args.no_e2e:
                ms3, nbytes, e2e_err = _time_e2e(tr2, x2, y2, xr2, s_level, 3, dev)
                e2e = {"value": nv2 / (ms3 * 1e-3) / 1e6, "unit": "Mvoxels/s", "workload": sw_name,
                       "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
                       "ms_per_step": ms3, "pr_rel_err": e2e_err,
                       "api": "per-rank pinned host slabs -> H2D, nddwt_mplan dec, D2H, H2D, rec, D2H (all ranks, max over ranks)"}
            del tr2, plan2, x2, y2, xr2
        except Exception as exc:   # noqa: BLE001
            same = {"workload": sw_name, "error": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": "dec+rec Mvoxels/s", "value": value, "unit": "Mvoxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c64" if dtype == "complex64" else "c128", "data": "synthetic",
            "config": {"workload": wl_name, "sizes": list(sizes), "wavelet": wname, "levels": level, "bands": nb,
                       "elem": dtype,
                       "parallelism": "slab%d (last dim, %s planes/GPU), halo planes pushed per level" % (
                           world, "/".join(str(c) for _, c in parts)),
                       "transport": ("peer memory: copy-engine pushes over NVLink into CUDA-IPC inboxes, flags in peer "
                                     "memory (nddwt_mplan_*)") if transport == "peer" else
                                    "NCCL send/recv per level (torch.distributed batch_isend_irecv)",
                       "transport_fallback": fallback, "flag_wait_timeouts": timeouts, "comm_streams": COMM_STREAMS or "default (1)", "z_chunks": Z_CHUNKS or "automatic (4 when the halo exceeds the slab, 2 down to a third of it, else 1)",
                       "l2": "per-GPU working set %.1f GB >> L2, no flush" % ((1 + nb) * nvox * esize / world / 1e9),
                       "pr_rel_err": pr_err, "dec_rel_err": e_dec, "parity_rec_rel_err": e_rec,
                       "parity_case": {"sizes": small, "what": "rank 0's slab of dec vs the oracle, reconstruction on every rank"},
                       "halo_bytes_per_rank_per_step": int(halo_bytes),
                       "dec_ms_rank0": dec_ms, "rec_ms_rank0": rec_ms, "rank0_kernel_times": ktimes, "overlap": overlap, "scatter_exchange": scatter,
                       "scaling_comparable": "N=1 runs cfg5 (cfg4 needs 197.6 GB of coefficients); "
                                             "`same_workload` is cfg5 sharded over the same ranks",
                       "same_workload": same},
            "roofline": {"bound": "hbm", "achieved": pair_gbs / world, "peak": peak, "unit": "GB/s",
                         "frac": pair_gbs / world / peak, "traffic": None,
                         "kernel": "whole dec+rec pair per GPU (compulsory bytes 2(1+nb)Ne / P)",
                         "peak_source": peak_src},
            "cpu_baseline": None,
            "e2e": e2e,
            "gpu_launches": int(launches) * world, "clocks": clocks,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
