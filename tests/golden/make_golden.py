"""Generates tests/golden/*.npz.

The reference (pure Julia) cannot be run in this image and stores no golden matrices of its own
(SURVEY.md section 4), so the fixtures are: (a) the reference's planar known answer, typed in
from test/usecases/simple.jl / README.md, and (b) matrices produced by the CPU oracle
(oracle/crg_oracle.c) on small grid pairs -- a regression pin for the oracle itself and a
GPU-box-portable checker for the CUDA path.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import GRID_PAIRS_SMALL  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    rng = np.random.default_rng(20260101)
    for name, (fd, fs) in GRID_PAIRS_SMALL.items():
        dst, src = fd(), fs()
        R = oracle.build_regridder(dst, src)
        x = rng.random(src.ncells)
        y = R.regrid(x)
        xb = R.regrid(y, transpose=True)
        fn = os.path.join(HERE, name.replace("<-", "__from__") + ".npz")
        np.savez_compressed(fn, colptr=R.colptr, rowval=R.rowval, nzval=R.nzval, dst_areas=R.dst_areas,
                            src_areas=R.src_areas, x=x, y=y, xb=xb, n_dst=R.n_dst, n_src=R.n_src)
        print(fn, R.nnz, os.path.getsize(fn))


if __name__ == "__main__":
    main()
