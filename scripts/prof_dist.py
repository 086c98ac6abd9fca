import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from crg_b200 import grids
from crg_b200.dist import ShardedRegridder, _LocalB200
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ws = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ws); stream = ws.cuda_stream
d = grids.lonlat_grid(1440, 720); s = grids.healpix_grid(512, "ring")
dd = grids.Grid(torch.from_numpy(d.verts).to(dev), d.manifold); sd = grids.Grid(torch.from_numpy(s.verts).to(dev), s.manifold)
x = torch.rand(s.ncells, dtype=torch.float64, device=dev)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0
def step():
    t = time.perf_counter(); dist.barrier(); tick("barrier", t)
    t = time.perf_counter(); S = ShardedRegridder(dd, sd, local_factory=lambda a, b: _LocalB200(a, b, stream=stream), device=dev); tick("build", t)
    t = time.perf_counter(); y = S.regrid(x if rank == 0 else None); tick("fwd", t)
    t = time.perf_counter(); xb = S.regrid(y, transpose=True); tick("T", t)
    return S
for _ in range(3): step()
T.clear()
for _ in range(10): S = step()
print(f"rank {rank}:", {k: round(v * 100, 3) for k, v in T.items()}, "build stats", {k: round(v, 2) for k, v in S.local.stats.items() if k in ("ms_total", "ms_device")}, flush=True)
dist.destroy_process_group()
