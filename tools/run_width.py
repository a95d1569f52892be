"""One SS2D_cond_v10 call at one of the UNet's other widths (the ncu target for the D = 1280 / 2560 launches).
    python tools/run_width.py 640|1280 [Bp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from actalker_b200 import SS2D_cond_v10

dm = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
Bp = int(sys.argv[2]) if len(sys.argv) > 2 else 50
side = int(72 / (dm / 320))
torch.manual_seed(72589)
layer = SS2D_cond_v10(d_model=dm, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=side, scan_type="sweep",
                      num_direction=2).eval().to(torch.bfloat16)
for n, p in layer.named_parameters():
    if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
        p.data = p.data.float()
layer = layer.cuda()
x = torch.randn(Bp, side * side, dm, device="cuda").to(torch.bfloat16)
idm = torch.randn(Bp, 1, 1024, device="cuda").to(torch.bfloat16)
cd = torch.randn(Bp, 33, 1024, device="cuda").to(torch.bfloat16)
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(4):
        y = layer(x, idm, cd, [ones, ones])
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.float().abs().mean()))
