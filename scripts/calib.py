import torch, numpy as np
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
def t(f, n=20):
    ts=[]
    for i in range(n):
        flush.sum()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return np.median(ts[5:])*1e3
for mb in (16, 111, 157, 512, 2048):
    a=torch.zeros(mb*(1<<20)//8, dtype=torch.float64, device="cuda"); b=torch.empty_like(a)
    us=t(lambda: a.sum()); print(f"sum  {mb} MB: {us:.1f} us -> {mb*1.048576/us*1e3:.0f} GB/s")
    us=t(lambda: b.copy_(a)); print(f"copy {mb} MB: {us:.1f} us -> {2*mb*1.048576/us*1e3:.0f} GB/s")
    us=t(lambda: a.mul_(1.0)) if mb<=512 else 0
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); e1.record(); torch.cuda.synchronize(); print("empty event pair", e0.elapsed_time(e1)*1e3, "us")
a=torch.zeros(8, device="cuda")
ts=[]
for i in range(20):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); a.add_(1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("tiny kernel", np.median(ts)*1e3, "us")
