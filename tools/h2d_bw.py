import torch, time
x=torch.empty(1_866_240_000, dtype=torch.uint8).pin_memory()
d=torch.empty_like(x, device='cuda')
for _ in range(2): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print("H2D pinned GB/s", x.numel()/dt/1e9)
s2=torch.cuda.Stream(); y=torch.empty_like(x).pin_memory(); d2=torch.empty_like(d)
torch.cuda.synchronize(); t=time.perf_counter()
d.copy_(x, non_blocking=True)
with torch.cuda.stream(s2): d2.copy_(y, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("2 concurrent H2D GB/s total", 2*x.numel()/dt/1e9)
t=time.perf_counter(); x2=d.cpu(); dt=time.perf_counter()-t; print("D2H pageable GB/s", x.numel()/dt/1e9)
import os; print("cpus", os.cpu_count())
