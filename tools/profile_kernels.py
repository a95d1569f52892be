#!/usr/bin/env python
"""Kernel-level time of one layer call (every kernel, torch glue included) from torch.profiler:
    python tools/profile_kernels.py [d_model] [B'] [ones|rects]
Prints the kernels by total device time over 5 calls — what to look at when the step is not scan-dominated."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bench_configs import make_layer

dm = int(sys.argv[1]) if len(sys.argv) > 1 else 320
Bp = int(sys.argv[2]) if len(sys.argv) > 2 else 50
kind = sys.argv[3] if len(sys.argv) > 3 else "rects"
dtype = torch.bfloat16
side = int(72 / (dm / 320)); L = side * side
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
mouth = torch.zeros_like(ones); mouth[:, :, 330:480, 180:400] = 1
upper = torch.zeros_like(ones); upper[:, :, 60:330, 100:480] = 1
masks = [ones, ones.clone()] if kind == "ones" else [mouth, upper]
layer = make_layer(dm)
x = torch.randn(Bp, L, dm, device="cuda").to(dtype)
idm = torch.randn(Bp, 1, 1024, device="cuda").to(dtype)
cd = torch.randn(Bp, 33, 1024, device="cuda").to(dtype)
with torch.no_grad():
    for _ in range(3):
        layer(x, idm, cd, masks)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            layer(x, idm, cd, masks)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"d_model {dm} B'={Bp} masks={kind}: {tot / 5 / 1e3:.3f} ms of kernels per call")
for e in rows[:16]:
    print(f"{e.device_time_total / 5:9.1f} us  x{e.count // 5:<3d} {e.key[:110]}")
