import torch, time
for mb in (8, 35, 139, 512):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(mb, "MB H2D", round(n * 10 / a.elapsed_time(b) / 1e6, 1), "GB/s")
    a.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(mb, "MB D2H", round(n * 10 / a.elapsed_time(b) / 1e6, 1), "GB/s")
