"""Build-only micro-benchmark (cfg5, device-resident vertices): per-phase device times, median of N."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids
from crg_b200.regridder import Regridder
d = grids.lonlat_grid(1440, 720); s = grids.healpix_grid(512, "ring")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
gd = grids.Grid(torch.from_numpy(d.verts).cuda(), d.manifold); gs = grids.Grid(torch.from_numpy(s.verts).cuda(), s.manifold)
rows = []; walls = []
for i in range(12):
    torch.cuda.synchronize(); t = time.perf_counter()
    R = Regridder(gd, gs, stream=st.cuda_stream)
    torch.cuda.synchronize(); w = (time.perf_counter() - t) * 1e3
    if i >= 4: rows.append(R.intersections.stats()); walls.append(w)
keys = [k for k in rows[0] if k.startswith("ms_")]
print("tag", os.environ.get("TAG", ""), "wall median %.2f ms" % np.median(walls), {k[3:]: round(float(np.median([r[k] for r in rows])), 3) for k in keys},
      "cand", rows[0]["n_candidates"], "nnz", rows[0]["nnz"])
