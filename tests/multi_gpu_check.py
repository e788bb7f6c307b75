#!/usr/bin/env python3
"""Multi-GPU parity check, one process per GPU (run under torchrun on a box with >= 2 GPUs; spawned by
tests/test_gpu_multi.py::test_one_process_per_gpu_torchrun): every rank transforms its slab of a seeded
global array and compares with the oracle's full-array transform, once per transport: the library's
peer-memory pushes (nddwt_mplan_*, CUDA IPC + flags) and the NCCL send/recv schedule (SlabTransform).

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nddwt_oracle as orc  # noqa: E402

slab = importlib.import_module("non-decimated_wavelets_b200.slab")
_lib = importlib.import_module("non-decimated_wavelets_b200._lib")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)   # NCCL kernels must not queue behind the tile kernels
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, pg_options=opts)
    worst = 0.0
    cases = [((32, 24, 16, 16), "db4", 3, 0), ((32, 24, 16, 4 * world), "db4", 2, 1), ((64, 32, 24), "db4", 2, 0),
             ((40, 20, 12, 2 * world + 1), "db2", 2, 0), ((32, 16, 8, 3 * world), "db1", 2, 1),
             ((32, 16, 32, 4 * world), "db4", 3, 0)]          # dim 3 long enough for the z-chunked (pipelined) exchange
    for sizes, wname, level, l2 in cases:
        d = len(sizes)
        wn = [wname] * d
        L = len(orc.wave_filters(wname)[0])
        x = orc.synth(sizes, np.complex64, 7)
        parts = slab.slab_partition(sizes[-1], world)
        s, c = parts[rank]
        xl = torch.from_numpy(np.ascontiguousarray(x[..., s:s + c].transpose(*reversed(range(d))))).to(dev)
        yo = orc.dec_direct(x.astype(np.complex128), wn, level, bool(l2))
        for transport in ("peer", "nccl"):
            if transport == "peer":
                tr = slab.PeerSlabTransform(sizes, wn, _lib.NDDWT_C64, l2, rank, world, lr)
                assert (tr.start, tr.n_local) == (s, c)
            else:
                eng = slab.CudaSlabEngine(tuple(sizes[:-1]) + (c,), sizes[-1], wn, _lib.NDDWT_C64, l2, lr)
                tr = slab.SlabTransform(sizes, wn, level, eng, L, rank, world, device=dev, dtype=torch.complex64)
            for rep in range(2):            # second pass: buffers, flags and sequence numbers are reused
                y = tr.dec(xl, level)
                xr = tr.rec(y)
            torch.cuda.synchronize()
            if transport == "peer":
                tr.plan.sync()
                assert tr.plan.wait_timeouts == 0
            y_np = y.cpu().numpy().transpose(*reversed(range(d + 1)))
            e_dec = orc.rel_l2(y_np, yo[..., s:s + c, :])
            e_rec = orc.rel_l2(xr.cpu().numpy().transpose(*reversed(range(d))), x[..., s:s + c])
            worst = max(worst, e_dec, e_rec)
            print("rank %d %s %s J%d %s overlap=%s: dec %.2e rec %.2e" % (rank, sizes, wname, level, transport,
                                                                          tr.overlap, e_dec, e_rec), flush=True)
            dist.barrier()
            del tr
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    if float(t[0]) > 1e-5:
        print("FAILED worst %.3e" % float(t[0]))
        sys.exit(1)
    if rank == 0:
        print("multi-GPU parity OK, worst rel-L2 %.2e" % float(t[0]))


if __name__ == "__main__":
    main()
