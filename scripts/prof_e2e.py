import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crg_b200 import grids, _lib
from crg_b200.regridder import Regridder, regrid_, transpose
ds, ss = grids.lonlat_spec(1440, 720), grids.healpix_spec(512, "ring")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = stream.cuda_stream
def pinned(a):
    t_ = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t_, t_.numpy()
xt, xh = pinned(np.random.rand(ss.ncells)); yt, yh = pinned(np.zeros(ds.ncells)); xbt, xbh = pinned(np.zeros(ss.ncells))
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0
def step():
    t = time.perf_counter(); R = Regridder(ds, ss, stream=sp); tick("build", t)
    t = time.perf_counter(); regrid_(yh, R, xh); tick("fwd", t)
    t = time.perf_counter(); regrid_(xbh, transpose(R), yh); tick("T", t)
    t = time.perf_counter(); a, b = R.dst_areas, R.src_areas; tick("areas", t)
    return R
for _ in range(3): R = step()
T.clear()
t0 = time.perf_counter()
for _ in range(10): R = step()
torch.cuda.synchronize(); print("e2e ms/step", (time.perf_counter() - t0) * 100, {k: round(v * 100, 3) for k, v in T.items()})
