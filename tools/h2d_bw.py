"""Pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e number): GB/s for a few sizes."""
import torch

for mb in (64, 400, 1600):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(f"{mb} MiB: {5 * n / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s")
