#!/usr/bin/env python3
"""Minimal driver for ncu: a few dec(+rec) calls of one workload (device-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nddwt_b200 as nd
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sizes, wname, level, dtype = WORKLOADS[name][:4]
batch = WORKLOADS[name][4] if len(WORKLOADS[name]) > 4 else 1
d = len(sizes)
cls = {1: nd.nd_dwt_1D, 2: nd.nd_dwt_2D, 3: nd.nd_dwt_3D, 4: nd.nd_dwt_4D}[d]
obj = cls(wname, list(sizes), "precision", "single", "compute", "gpu")
g = torch.Generator(device="cuda").manual_seed(0)
shape = tuple(sizes) + ((batch,) if batch > 1 else ())   # batch extension: x is [sizes, B], column-major
x = torch.view_as_complex(torch.randn(tuple(reversed(shape)) + (2,), generator=g, device="cuda")).permute(*reversed(range(len(shape))))
for _ in range(reps):
    y = obj.dec(x, level)
    xr = obj.rec(y)
torch.cuda.synchronize()
print("ok", float(torch.linalg.vector_norm(xr - x) / torch.linalg.vector_norm(x)))
