"""Profile driver: one cfg5 build, then forward / transpose SpMV launches (for ncu -k regex:spmv)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid_, transpose
small = "--small" in sys.argv
d = grids.lonlat_grid(720, 360) if small else grids.lonlat_grid(1440, 720)
s = grids.healpix_grid(256, "ring") if small else grids.healpix_grid(512, "ring")
gd = grids.Grid(torch.from_numpy(d.verts).cuda(), d.manifold); gs = grids.Grid(torch.from_numpy(s.verts).cuda(), s.manifold)
R = Regridder(gd, gs)
x = torch.rand(s.ncells, dtype=torch.float64, device="cuda"); y = torch.zeros(d.ncells, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_(); regrid_(y, R, x); flush.zero_(); regrid_(x, transpose(R), y)
torch.cuda.synchronize()
