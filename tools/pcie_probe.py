"""Host<->device copy bandwidth from pinned memory for the sizes the pose batches use."""
import torch, time
torch.cuda.init()
for mb in (0.2, 1.7, 3.4, 13.6, 54.5, 218):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    h2d = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    e0.record()
    for _ in range(reps): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    d2h = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"{mb:7.1f} MB  H2D {h2d:6.1f} GB/s   D2H {d2h:6.1f} GB/s")
