#!/usr/bin/env python
"""raw pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e number): python tools/h2d_bw.py"""
import time

import torch

n = 65536 * 832
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D {n * 4 / 1e6:.0f} MB pinned: {dt * 1e3:.2f} ms -> {n * 4 / dt / 1e9:.1f} GB/s = {65536 / dt / 1e6:.2f} M BG2 codewords/s")
