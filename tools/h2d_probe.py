"""host->device copy rate of this box for the e2e leg's transfer sizes (pinned memory, one stream)"""
import torch, time
for mb in (1, 6.3, 12.6, 64, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb:6.1f} MB: {ms*1e3:8.1f} us  {n/ms/1e6:6.1f} GB/s")
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"D2H {mb:6.1f} MB: {ms*1e3:8.1f} us  {n/ms/1e6:6.1f} GB/s")
