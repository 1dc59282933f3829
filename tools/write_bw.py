"""Development probe: pure write / read / copy bandwidth at the sizes the refit kernel moves (CUDA events)."""
import torch
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for mb in (48, 192, 1024):
    n = mb * (1 << 20) // 4
    a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    w = t(lambda: a.zero_()); r = t(lambda: a.sum()); c = t(lambda: b.copy_(a))
    print(f"{mb:5d} MB: write(zero_) {w:7.1f} us = {mb*1.048576/w*1e3:6.0f} GB/s | read(sum) {r:7.1f} us = {mb*1.048576/r*1e3:6.0f} GB/s | copy {c:7.1f} us = {2*mb*1.048576/c*1e3:6.0f} GB/s (r+w)")
