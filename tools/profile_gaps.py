#!/usr/bin/env python
"""Device timeline of one layer call at BASELINE configs[1]: every kernel / memset with its start offset, duration and the
idle gap before it (torch.profiler / CUPTI timestamps) — where the step's time goes BETWEEN the kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bench_configs import make_layer

Bp = int(sys.argv[1]) if len(sys.argv) > 1 else 25
dtype = torch.bfloat16
layer = make_layer(320)
L = 5184
xs = [torch.randn(Bp, L, 320, device="cuda").to(dtype) for _ in range(3)]
idm = torch.randn(Bp, 1, 1024, device="cuda").to(dtype)
cd = torch.randn(Bp, 33, 1024, device="cuda").to(dtype)
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
masks = [ones, ones.clone()]
with torch.no_grad():
    for i in range(6):
        layer(xs[i % 3], idm, cd, masks)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(4):
            layer(xs[i % 3], idm, cd, masks)
        torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_time_total > 0 or "Mem" in e.name), key=lambda e: e.time_range.start)
t0 = None
prev_end = None
n = len(ev) // 4
for e in ev[2 * n:3 * n + 1]:      # the third call and the first kernel of the fourth
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    if t0 is None:
        t0 = s
    gap = 0.0 if prev_end is None else s - prev_end
    print(f"t={s - t0:9.1f} us  gap {gap:6.1f}  dur {d:8.1f}  {e.name[:90]}")
    prev_end = e.time_range.end
