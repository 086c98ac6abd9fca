"""0.125 deg lon-lat <-> HEALPix 1024 (4x BASELINE config 5 in every count): steady-state build phases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crg_b200 import grids
from crg_b200.regridder import Regridder
for i in range(4):
    R = Regridder(grids.lonlat_spec(2880, 1440), grids.healpix_spec(1024, "ring"))
    st = R.intersections.stats()
    print(i, {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")}, st["n_candidates"], R.intersections.nnz, flush=True)
    del R
