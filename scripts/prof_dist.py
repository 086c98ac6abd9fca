"""Host-side profile of one sharded step (run under torchrun): where the milliseconds of build / regrid! / transpose go."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from crg_b200 import grids
from crg_b200.dist import ShardedRegridder, _LocalB200
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ws = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ws); stream = ws.cuda_stream
d = grids.lonlat_spec(1440, 720); s = grids.healpix_spec(512, "ring")
x = torch.rand(s.ncells, dtype=torch.float64, device=dev)
T = {}
def tick(name, t0, sync=True):
    if sync: torch.cuda.synchronize()
    T[name] = T.get(name, 0) + time.perf_counter() - t0
def step(sync=True):
    t = time.perf_counter(); dist.barrier(); tick("barrier", t)
    t = time.perf_counter(); S = ShardedRegridder(d, s, local_factory=lambda a, b: _LocalB200(a, b, stream=stream), device=dev); tick("build", t, sync)
    t = time.perf_counter(); y = S.regrid(x if rank == 0 else None); tick("fwd", t, sync)
    t = time.perf_counter(); xb = S.regrid(y, transpose=True); tick("T", t, sync)
    return S
for _ in range(3): step()
for sync in (True, False):
    T.clear()
    for _ in range(20): S = step(sync)
    torch.cuda.synchronize()
    print(f"rank {rank} sync={sync} ms/step:", {k: round(v * 50, 3) for k, v in T.items()}, "build stats", {k: round(v, 2) for k, v in S.local.stats.items() if k in ("ms_total", "ms_device")}, flush=True)
if rank == 0:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20): step(False)
    pr.disable(); torch.cuda.synchronize()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28); print(st.getvalue()[:4500])
else:
    for _ in range(20): step(False)
dist.destroy_process_group()
