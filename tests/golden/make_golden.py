#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the oracle (oracle/nddwt_oracle.py, the restated 'mat' FFT
path of the reference).  Run in the build container:  python tests/golden/make_golden.py

The reference holds no golden vectors of its own; where /root/reference and
oracle/_ref are available the script also checks every double-precision case against the
reference's compiled native core before writing, so the fixtures are outputs the reference's own
C code reproduces.  Inputs are seeded (oracle.synth) and stored with the outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nddwt_oracle as orc  # noqa: E402
from oracle import ref_mex  # noqa: E402

CASES = [
    # name, sizes, wname, level, pres_l2, dtype
    ("d1_db1_J4", (61,), "db1", 4, 0, "complex128"),          # mini Test/nddwt1D_test.m:5-8
    ("d1_db8_J3_f32", (64,), "db8", 3, 1, "float32"),
    ("d1_db10_J2", (47,), "db10", 2, 0, "float64"),
    ("d2_mixed_J1", (24, 20), ["db1", "db3"], 1, 1, "complex128"),   # mini Test/nddwt2D_test.m:5-8
    ("d2_db4_J3_odd_c64", (17, 19), "db4", 3, 0, "complex64"),
    ("d3_mixed_J2", (12, 10, 9), ["db1", "db3", "db1"], 2, 1, "complex128"),  # mini Test/nddwt3D_test.m
    ("d3_db4_J3_c64", (12, 10, 12), "db4", 3, 0, "complex64"),
    ("d4_mixed_J1", (8, 6, 4, 6), ["db1", "db3", "db1", "db1"], 1, 1, "complex128"),  # mini Test/nddwt4D_test.m
    ("d4_db2_J2_c64", (6, 6, 4, 6), "db2", 2, 0, "complex64"),
    ("d2_haar_J1", (10, 14), "db1", 1, 1, "float64"),
]


def main():
    have_ref = ref_mex.available()
    for seed, (name, sizes, wname, level, l2, dtype) in enumerate(CASES):
        x = orc.synth(sizes, dtype, 100 + seed)
        prec = "single" if np.dtype(dtype).itemsize in (4, 8) and np.dtype(dtype) in (np.float32, np.complex64) else "double"
        y = orc.dec(x, wname, level, bool(l2), precision=prec)
        xr = orc.rec(y, wname, bool(l2), precision=prec)
        tol = 1e-5 if prec == "single" else 1e-12
        assert orc.rel_l2(xr, x) < tol, name
        if have_ref and prec == "double":
            yref = ref_mex.dec(x, wname, level, bool(l2))
            assert orc.rel_l2(yref, y) < 1e-13, (name, orc.rel_l2(yref, y))
        yd = orc.dec_direct(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), wname, level, bool(l2))
        assert orc.rel_l2(yd, y) < tol, name
        np.savez(os.path.join(HERE, name + ".npz"), x=x, y=y.astype(x.dtype),
                 wname=np.array(wname if isinstance(wname, list) else [wname]),
                 level=level, pres_l2=l2)
        print(name, y.shape, y.dtype, "ref-checked" if (have_ref and prec == "double") else "")


if __name__ == "__main__":
    main()
