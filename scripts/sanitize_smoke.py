"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid_, transpose, regridder_from_coo
# spherical quads (fast path), described grids, cut slices (long polar rows), ragged planar polygons, from_coo
R = Regridder(grids.healpix_spec(2, "ring"), grids.lonlat_spec(720, 30, 0, 360, 75, 90))
x = np.random.rand(R.shape[1]); y = np.zeros(R.shape[0]); regrid_(y, R, x); xb = np.zeros(R.shape[1]); regrid_(xb, transpose(R), y)
X = np.random.rand(R.shape[1], 5); Y = np.zeros((R.shape[0], 5)); regrid_(Y, R, X); regrid_(np.asfortranarray(Y), R, np.asfortranarray(X))
R2 = Regridder(grids.lonlat_grid(36, 18), grids.healpix_grid(8, "nested"), normalize=True)
y2 = np.zeros(R2.shape[0]); regrid_(y2, R2, np.random.rand(R2.shape[1]))
rng = np.random.default_rng(0)
def poly(k):
    t = np.sort(rng.random(k)) * 2 * np.pi; c = rng.random(2) * 3
    return np.stack([c[0] + 0.5 * np.cos(t), c[1] + 0.5 * np.sin(t)], 1)
R3 = Regridder(grids.polygons_grid([poly(rng.integers(3, 9)) for _ in range(60)]), grids.polygons_grid([poly(rng.integers(3, 7)) for _ in range(80)]))
R4 = regridder_from_coo(50, 40, rng.integers(0, 50, 500), rng.integers(0, 40, 500), rng.random(500) + 0.1, np.ones(50), np.ones(40))
y4 = np.zeros(50); regrid_(y4, R4, np.ones(40))
# identical cells (coincident edges: snapped crossings, sequential fallback), a destination block (culling box)
R5 = Regridder(grids.healpix_grid(8, "nested"), grids.healpix_grid(8, "ring"))
d6 = grids.lonlat_grid(48, 24)
R6 = Regridder(d6.slice(200, 500), grids.cubed_sphere_grid(8), build_transpose=True)
print("sanitize smoke done", R.intersections.nnz, R2.intersections.nnz, R3.intersections.nnz, R4.intersections.nnz)
