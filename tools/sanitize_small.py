#!/usr/bin/env python3
"""Small fused + generic cases for compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nddwt_b200 as nd
from oracle import nddwt_oracle as orc
for sizes, wn, lv in [((64, 48, 20), "db4", 2), ((40, 36, 12, 8), "db4", 2), ((34, 18, 10), "db2", 1),
                      ((32, 20, 9, 5), "db1", 2), ((33, 17), "db3", 2), ((100,), "db8", 3), ((48, 40, 8, 6), "db3", 1)]:
    cls = {1: nd.nd_dwt_1D, 2: nd.nd_dwt_2D, 3: nd.nd_dwt_3D, 4: nd.nd_dwt_4D}[len(sizes)]
    x = orc.synth(sizes, np.complex64, 1)
    o = cls(wn, list(sizes), "precision", "single")
    y = o.dec(x, lv)
    xr = o.rec(y)
    print(sizes, wn, lv, "PR err %.2e" % orc.rel_l2(xr, x), flush=True)
print("done")
