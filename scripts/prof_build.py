"""Profile driver: cfg5 (or cfg2 with --small) build x N + forward/transpose apply."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids
from crg_b200.regridder import Regridder, regrid_, transpose
small = "--small" in sys.argv
reps = 2
d = grids.lonlat_grid(720, 360) if small else grids.lonlat_grid(1440, 720)
s = grids.healpix_grid(256, "ring") if small else grids.healpix_grid(512, "ring")
dv = torch.from_numpy(d.verts).cuda(); sv = torch.from_numpy(s.verts).cuda()
gd = grids.Grid(dv, d.manifold); gs = grids.Grid(sv, s.manifold)
for r in range(reps):
    t = time.time(); R = Regridder(gd, gs); torch.cuda.synchronize(); w = time.time() - t
    st = R.intersections.stats()
    print("build wall %.2f ms" % (w * 1e3), {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")}, flush=True)
x = torch.rand(s.ncells, dtype=torch.float64, device="cuda"); y = torch.zeros(d.ncells, dtype=torch.float64, device="cuda")
regrid_(y, R, x); regrid_(x, transpose(R), y)
X = torch.rand(s.ncells, 16, dtype=torch.float64, device="cuda"); Y = torch.zeros(d.ncells, 16, dtype=torch.float64, device="cuda")
regrid_(Y, R, X)
regrid_(Y.T.contiguous().T, R, X.T.contiguous().T)
torch.cuda.synchronize()
