"""Timing of the layer's dense GEMMs as cuBLAS runs them (config 2 shapes), to decide what is worth restructuring."""
import torch
torch.manual_seed(0)
dev = "cuda"
M, dm, D, xw = 25 * 5184, 320, 640, 128
dt = torch.bfloat16
x = torch.randn(M, dm, device=dev, dtype=dt)
W1 = torch.randn(D, dm, device=dev, dtype=dt); W2 = torch.randn(D, dm, device=dev, dtype=dt)
Wc = torch.cat([W1, W2], 0).contiguous()
Wx = torch.randn(xw, D, device=dev, dtype=dt)
Wdt = torch.randn(64, 2 * D, device=dev, dtype=dt)
Wo = torch.randn(dm, D, device=dev, dtype=dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def t(fn, n=20):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1e3

xz = torch.nn.functional.linear(x, Wc)            # (M, 2D)
xz1 = torch.nn.functional.linear(x, W1)
xd = torch.nn.functional.linear(xz1, Wx)
y = torch.randn(M, D, device=dev, dtype=dt)
print("in_proj x2 separate      %.1f us" % t(lambda: (torch.nn.functional.linear(x, W1), torch.nn.functional.linear(x, W2))))
print("in_proj stacked N=1280   %.1f us" % t(lambda: torch.nn.functional.linear(x, Wc)))
print("x_proj dense input       %.1f us" % t(lambda: torch.mm(xz1, Wx.t())))
print("x_proj strided input     %.1f us" % t(lambda: torch.mm(xz[:, :D], Wx.t())))
print("dt_proj (lda=xw view)    %.1f us" % t(lambda: torch.mm(xd[:, 64:], Wdt)))
print("out_proj                 %.1f us" % t(lambda: torch.nn.functional.linear(y, Wo)))
Wbd = torch.zeros(2 * xw, 2 * D, device=dev, dtype=dt); Wbd[:xw, :D] = Wx; Wbd[xw:, D:] = Wx
print("x_proj both (blockdiag)  %.1f us" % t(lambda: torch.mm(xz, Wbd.t())))
Wb = torch.stack([W1.t().contiguous(), W2.t().contiguous()], 0)          # (2, K, N)
xe = x.unsqueeze(0).expand(2, M, dm)
print("in_proj bmm, expanded x  %.1f us" % t(lambda: torch.bmm(xe, Wb)))
outb = torch.empty(2, M, D, device=dev, dtype=dt)
print("in_proj bmm out=         %.1f us" % t(lambda: torch.bmm(xe, Wb, out=outb)))
print("in_proj matmul broadcast %.1f us" % t(lambda: torch.matmul(x.unsqueeze(0), Wb)))
