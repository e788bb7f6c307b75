#!/usr/bin/env python3
"""Quick GPU check of a tuning variant (NDDWT_VARIANT etc. from the environment): fused kernels vs the
generic kernels (kernel_mode=1) and the oracle on a few small complex-single shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nddwt_b200 as nd
from oracle import nddwt_oracle as orc

CASES = [((64, 48, 40), "db4", 3), ((32, 32, 24, 16), "db4", 2), ((192, 40, 16, 8), "db4", 1),
         ((64, 30, 9, 8), "db4", 1), ((128, 17, 12), "db4", 1), ((256, 24, 10, 8), "db4", 1), ((72, 20, 8, 8), "db4", 1)]
worst = 0.0
for sizes, wn, level in CASES:
    x = orc.synth(sizes, np.complex64, 3)
    cls = {3: nd.nd_dwt_3D, 4: nd.nd_dwt_4D}[len(sizes)]
    a = cls(wn, list(sizes), "precision", "single", "compute", "mex")
    b = cls(wn, list(sizes), "precision", "single", "compute", "mex")
    b.set_kernel_mode(1)
    ya, yb = a.dec(x, level), b.dec(x, level)
    c = orc.synth(ya.shape, np.complex64, 4)
    e = [orc.rel_l2(ya, yb), orc.rel_l2(a.rec(ya), x), orc.rel_l2(a.rec(c), b.rec(c))]
    worst = max(worst, *e)
    print(sizes, ["%.2e" % v for v in e], "synthesis kernel", a.synthesis_kernels())
print("worst", "%.3e" % worst, "OK" if worst < 2e-6 else "FAIL")
sys.exit(0 if worst < 2e-6 else 1)
