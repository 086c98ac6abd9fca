import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for mb in (1, 8, 25, 256):
    x = torch.ones(mb * (1 << 20) // 8, dtype=torch.float64, device=dev)
    for _ in range(5): dist.all_reduce(x); dist.broadcast(x, 0)
    torch.cuda.synchronize(); dist.barrier()
    for name, f in (("all_reduce", lambda: dist.all_reduce(x)), ("broadcast", lambda: dist.broadcast(x, 0))):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        if rank == 0: print(f"{name} {mb} MB: {e0.elapsed_time(e1)/10*1e3:.1f} us", flush=True)
if rank == 0:
    print("can_device_access_peer", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
