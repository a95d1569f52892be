"""One config-3 style call with the mouth / upper-face rectangle masks (d_model 320, B'=50) — the ncu target for the
ragged-tile path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_configs import make_layer, timed

dtype = torch.bfloat16
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
mouth = torch.zeros_like(ones); mouth[:, :, 330:480, 180:400] = 1
upper = torch.zeros_like(ones); upper[:, :, 60:330, 100:480] = 1
layer = make_layer(320)
Bp, L = int(sys.argv[1]) if len(sys.argv) > 1 else 50, 5184
x = torch.randn(Bp, L, 320, device="cuda").to(dtype)
idm = torch.randn(Bp, 1, 1024, device="cuda").to(dtype)
cd = torch.randn(Bp, 33, 1024, device="cuda").to(dtype)
with torch.no_grad():
    ms, scan = timed(lambda: layer(x, idm, cd, [mouth, upper]))
print(f"rect masks B'={Bp}: layer {ms:.3f} ms, scan {scan:.3f} ms")
