"""Debug: crg_clip_pairs vs the build on the same candidate list; prints the pairs that differ with the host classification."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from crg_b200 import grids
from crg_b200.regridder import Regridder, clip_pairs
for dst, src in [(grids.lonlat_grid(90, 45), grids.healpix_grid(16, "nested")), (grids.lonlat_grid(360, 180), grids.healpix_grid(64, "ring"))]:
    R = Regridder(dst, src, keep_candidates=True)
    ps, pd = R.intersections.candidates()
    a = clip_pairs(dst, src, ps, pd)
    a2 = clip_pairs(dst, src, ps, pd)
    A = R.intersections.tocsr()
    b = np.asarray(A[pd, ps]).ravel()
    bad = np.where(a != b)[0]
    print("pairs", len(ps), "differ", len(bad), "clip_pairs repeatable", np.array_equal(a, a2), "nnz", A.nnz, (a > 0).sum())
    R2 = Regridder(dst, src, keep_candidates=True)
    print("build repeatable", (R2.intersections.tocsr() != A).nnz == 0)
    for k in bad[:12]:
        print(k, k % 32, ps[k], pd[k], a[k], b[k])
