"""Pinned host->device bandwidth on the box (explains the e2e figure of bench.py: lm_commit moves 0.5 GiB per commit)."""
import time
import torch

for mib in (64, 512):
    n = mib << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"H2D pinned {mib} MiB: {ms:.2f} ms  {n / ms / 1e6:.1f} GB/s")
    e0.record()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"D2H pinned {mib} MiB: {ms:.2f} ms  {n / ms / 1e6:.1f} GB/s")
