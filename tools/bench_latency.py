#!/usr/bin/env python
"""Small-batch latency of one layer call, eager vs CUDA-graph replay."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from actalker_b200 import SS2D_cond_v10
from actalker_b200.graphed import GraphedLayer

def layer_for(dm):
    torch.manual_seed(0)
    l = SS2D_cond_v10(d_model=dm, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=int(72 / (dm / 320)),
                      scan_type="sweep", num_direction=2).eval().to(torch.bfloat16)
    for n, p in l.named_parameters():
        if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
            p.data = p.data.float()
    return l.cuda()

def t(fn, it=50):
    for _ in range(5): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it

for dm, Bp in ((320, 1), (320, 4), (640, 4), (1280, 4), (1280, 100)):
    layer = layer_for(dm); side = int(72 / (dm / 320)); L = side * side
    x = torch.randn(Bp, L, dm, device="cuda").bfloat16(); i = torch.randn(Bp, 1, 1024, device="cuda").bfloat16()
    c = torch.randn(Bp, 33, 1024, device="cuda").bfloat16(); ones = torch.ones(1, 1, 576, 576, device="cuda").bfloat16()
    with torch.no_grad():
        eager = t(lambda: layer(x, i, c, [ones, ones]))
    g = GraphedLayer(layer, x, i, c, [ones, ones])
    graph = t(lambda: g(x, i, c))
    print(json.dumps({"d_model": dm, "Bp": Bp, "L": L, "eager_ms": round(eager, 4), "graph_ms": round(graph, 4)}))
