import torch, time
x = torch.empty(512<<20, dtype=torch.uint8).pin_memory()
d = torch.empty(512<<20, dtype=torch.uint8, device="cuda")
for n in (1,):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(4): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("H2D GB/s", 4*0.5368/dt)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(4): x.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("D2H GB/s", 4*0.5368/dt)
import os
print(os.cpu_count())
