"""Where does the end-to-end (host-buffer) step lose time against the device-resident one?  Times the compute stream's
layer calls inside the streamed pipeline and the gaps between them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_configs import make_layer
from actalker_b200.host_api import HostStreamedLayer

dtype = torch.bfloat16
layer = make_layer(320)
Bp, L = 25, 5184
hx = torch.randn(Bp, L, 320).to(dtype).pin_memory(); hid = torch.randn(Bp, 1, 1024).to(dtype).pin_memory()
hcd = torch.randn(Bp, 33, 1024).to(dtype).pin_memory()
hys = [torch.empty(Bp, L, 320, dtype=dtype).pin_memory() for _ in range(2)]
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
masks = [ones, ones.clone()]
runner = HostStreamedLayer(layer)
orig = runner.layer
marks = []
class Timed:
    def __call__(self, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(torch.cuda.current_stream()); y = orig(*a); e.record(torch.cuda.current_stream())
        marks.append((s, e)); return y
    def parameters(self): return orig.parameters()
runner.layer = Timed()
for i in range(5): runner.submit(hx, hid, hcd, masks, hys[i % 2])
runner.drain(); marks.clear()
n = 20
for i in range(n): runner.submit(hx, hid, hcd, masks, hys[i % 2])
runner.drain()
dur = [s.elapsed_time(e) for s, e in marks]
gap = [marks[i][1].elapsed_time(marks[i + 1][0]) for i in range(n - 1)]
print("layer call on the compute stream: mean %.3f ms (min %.3f max %.3f); gap between calls: mean %.3f ms (max %.3f)" %
      (sum(dur) / n, min(dur), max(dur), sum(gap) / len(gap), max(gap)))
# device-resident reference
x, idm, cd = hx.cuda(), hid.cuda(), hcd.cuda()
with torch.no_grad():
    for _ in range(5): orig(x, idm, cd, masks)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): orig(x, idm, cd, masks)
    e.record(); torch.cuda.synchronize()
print("device-resident: %.3f ms per call" % (s.elapsed_time(e) / n))
