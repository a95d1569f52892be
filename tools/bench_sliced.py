"""Scan time of one rank's d_inner slice (channel sharding emulated on one GPU) under the launch-shape heuristics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_configs import make_layer, timed
from actalker_b200 import mamba_layer as ml

dtype = torch.bfloat16
layer = make_layer(320)
Bp, L = 25, 5184
x = torch.randn(Bp, L, 320, device="cuda").to(dtype)
idm = torch.randn(Bp, 1, 1024, device="cuda").to(dtype)
cd = torch.randn(Bp, 33, 1024, device="cuda").to(dtype)
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
with torch.no_grad():
    proj = layer.project_inputs(x, idm, cd, [ones, ones])
    for parts in (2, 4, 8):
        w = 640 // parts
        for seg, chain in ((None, None), (1, 0), (1, 8), (1, 16), (2, 0), (3, 0)):
            ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = seg, chain
            ms, scan = timed(lambda: layer.scan_core(*proj, ch_slice=(0, w)))
            print(f"slice 1/{parts} ({w} ch): segments={seg} chain={chain}: core {ms:.3f} ms, scan {scan:.3f} ms", flush=True)
ml.SCAN_SEGMENTS = ml.SCAN_CHAIN = None
