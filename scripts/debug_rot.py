import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from crg_b200 import grids
from crg_b200.regridder import Regridder, clip_pairs
from oracle import oracle
from test_gpu_build import _rotation
rng = np.random.default_rng(42)
Q1, Q2 = _rotation(rng), _rotation(rng)
dst = grids.Grid(np.ascontiguousarray(grids.healpix_grid(16, "ring").verts @ Q1.T), grids.SPHERICAL)
src = grids.Grid(np.ascontiguousarray(grids.lonlat_grid(72, 36).verts @ Q2.T), grids.SPHERICAL)
print("single pair:", clip_pairs(dst, src, [2520], [462]), "oracle", oracle.intersection_area(1, src.cell(2520), dst.cell(462)))
one_s = grids.Grid(src.verts[2520:2521].copy(), 1); one_d = grids.Grid(dst.verts[462:463].copy(), 1)
print("isolated cells:", clip_pairs(one_d, one_s, [0], [0]))
# south-pole analogue
ps, pd = oracle.candidate_pairs_safe(dst, src)
a = clip_pairs(dst, src, ps, pd)
i1, i2, oa = oracle.compute_intersection_areas(dst, src, ps, pd)
key = pd * src.ncells + ps; order = np.argsort(key)
want = np.zeros(len(ps)); want[order[np.searchsorted(key[order], i2 * src.ncells + i1)]] = oa
bad = np.abs(a - want) > 1e-13
print("bad", bad.sum(), "rows of bad src:", np.unique(ps[bad] // 72), "n cuts? ")
# swap roles: dst = lonlat (polar cells as clip), src = healpix
ps2, pd2 = oracle.candidate_pairs_safe(src, dst)
a2 = clip_pairs(src, dst, ps2, pd2)
i1, i2, oa = oracle.compute_intersection_areas(src, dst, ps2, pd2)
key = pd2 * dst.ncells + ps2; order = np.argsort(key)
want2 = np.zeros(len(ps2)); want2[order[np.searchsorted(key[order], i2 * dst.ncells + i1)]] = oa
print("swapped roles bad", (np.abs(a2 - want2) > 1e-13).sum())
# unrotated
d0, s0 = grids.healpix_grid(16, "ring"), grids.lonlat_grid(72, 36)
ps3, pd3 = oracle.candidate_pairs_safe(d0, s0)
a3 = clip_pairs(d0, s0, ps3, pd3)
i1, i2, oa = oracle.compute_intersection_areas(d0, s0, ps3, pd3)
key = pd3 * s0.ncells + ps3; order = np.argsort(key)
want3 = np.zeros(len(ps3)); want3[order[np.searchsorted(key[order], i2 * s0.ncells + i1)]] = oa
b3 = np.abs(a3 - want3) > 1e-13
print("unrotated bad", b3.sum(), np.unique(ps3[b3] // 72))
