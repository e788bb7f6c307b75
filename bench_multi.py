"""bench_multi.py -- the N>1 leg of bench.py: one process per GPU (torchrun), slabs along the last
dimension, NCCL halo exchange per level.  Strong scaling on the named workload (cfg4 by default)."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def run_multi(args, wl_name, wl):
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, peaks
    slab = importlib.import_module("non-decimated_wavelets_b200.slab")
    _lib = importlib.import_module("non-decimated_wavelets_b200._lib")
    import nddwt_b200 as nd

    sizes, wname, level, dtype = wl
    d = len(sizes)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)   # NCCL kernels must not queue behind the tile kernels
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, pg_options=opts)

    wn = [wname] * d
    L = len(nd.wave_filters(wname)[0])
    parts = slab.slab_partition(sizes[-1], world)
    s0, c0 = parts[rank]
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    code = _lib.NDDWT_C64 if dtype == "complex64" else _lib.NDDWT_C128
    eng = slab.CudaSlabEngine(tuple(sizes[:-1]) + (c0,), sizes[-1], wn, code, 0, local_rank)
    tr = slab.SlabTransform(sizes, wn, level, eng, L, rank, world, device=dev, dtype=tdt)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    rdt = torch.float32 if dtype == "complex64" else torch.float64
    x = torch.view_as_complex(torch.randn(tr.local_shape + (2,), generator=g, device=dev, dtype=rdt))
    nb = tr.num_bands(level)
    y = torch.empty((nb,) + tr.local_shape, dtype=tdt, device=dev)
    xr = torch.empty_like(x)

    split = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step(timed=False):
        if timed:
            split[0].record()
        tr.dec(x, level, out=y)
        if timed:
            split[1].record()
        tr.rec(y, out=xr)
        if timed:
            split[2].record()

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    err = torch.linalg.vector_norm(xr - x) ** 2
    nrm = torch.linalg.vector_norm(x) ** 2
    both = torch.stack([err, nrm])
    dist.all_reduce(both)
    pr_err = float(torch.sqrt(both[0] / both[1]))

    dist.barrier()
    torch.cuda.synchronize()
    step(timed=True)
    torch.cuda.synchronize()
    dec_ms, rec_ms = split[0].elapsed_time(split[1]), split[1].elapsed_time(split[2])
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.plan.launches
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    clocks = sampler.stop() if sampler else None
    launches = eng.plan.launches - l0
    ms_per_step = float(ms[0]) / args.steps
    nvox = int(np.prod(sizes))
    esize = np.dtype(dtype).itemsize
    value = nvox / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = peaks()
    pair_bytes = 2 * (1 + nb) * nvox * esize
    pair_gbs = pair_bytes / (ms_per_step * 1e-3) / 1e9
    plane_bytes = int(np.prod(sizes[:-1])) * esize
    halo_bytes = level * (L - 1) * plane_bytes * (2 if getattr(tr, 'scatter', False) else 3)   # per rank per pair
    if rank == 0:
        line = {
            "metric": "dec+rec Mvoxels/s", "value": value, "unit": "Mvoxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c64" if dtype == "complex64" else "c128", "data": "synthetic",
            "config": {"workload": wl_name, "sizes": list(sizes), "wavelet": wname, "levels": level, "bands": nb,
                       "elem": dtype, "parallelism": "slab%d (last dim, %s planes/GPU), NCCL halo exchange per level"
                       % (world, "/".join(str(c) for _, c in parts)),
                       "l2": "per-GPU working set %.1f GB >> L2, no flush" % ((1 + nb) * nvox * esize / world / 1e9),
                       "pr_rel_err": pr_err, "halo_bytes_per_rank_per_step": halo_bytes,
                       "dec_ms_rank0": dec_ms, "rec_ms_rank0": rec_ms, "overlap": bool(tr.overlap),
                       "scatter_exchange": bool(getattr(tr, "scatter", False))},
            "roofline": {"bound": "hbm", "achieved": pair_gbs / world, "peak": peak, "unit": "GB/s",
                         "frac": pair_gbs / world / peak, "traffic": None,
                         "kernel": "whole dec+rec pair per GPU (compulsory bytes 2(1+nb)Ne / P)",
                         "peak_source": peak_src},
            "cpu_baseline": None,
            "e2e": {"value": None, "unit": "Mvoxels/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "unavailable": "the cfg4 coefficient stack is 197.6 GB: host buffers for it do not fit the box's "
                                   "196 GB of RAM at any N; the host-buffer path is measured at N=1 (cfg5)"},
            "gpu_launches": int(launches) * world, "clocks": clocks,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
