"""torchrun --nproc-per-node N tools/check_p2p.py — channel-sharded layer with the fused NVLink push gather against the
NCCL all-gather route (must be bit-identical: same kernels, same values, different transport) and against the
unsharded layer; then times both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tools.bench_configs import make_layer
from actalker_b200.sharded import ShardedSS2DCondV10

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dtype = torch.bfloat16
layer = make_layer(320)
Bp, L = 25, 5184
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.randn(Bp, L, 320, device="cuda", generator=g).to(dtype)
idm = torch.randn(Bp, 1, 1024, device="cuda", generator=g).to(dtype)
cd = torch.randn(Bp, 33, 1024, device="cuda", generator=g).to(dtype)
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
rect = torch.zeros_like(ones); rect[:, :, 60:330, 100:480] = 1
nccl = ShardedSS2DCondV10(layer, mode="channel", gather="nccl")
p2p = ShardedSS2DCondV10(layer, mode="channel", gather="p2p")
ok = True
with torch.no_grad():
    for masks in ([ones, ones], [ones, rect]):
        want = layer(x, idm, cd, masks)
        a = nccl(x, idm, cd, masks)
        for it in range(4):                       # several calls: exercises the double-buffered gather buffers
            b = p2p(x + 0.0 * it, idm, cd, masks)
            ok &= bool(torch.equal(a, b))
        err = (a.float() - want.float()).abs().max().item()
        ok &= err < 0.1
        if rank == 0:
            print(f"masks {'ones' if masks[1] is ones else 'rect'}: p2p == nccl: {torch.equal(a, b)}, max|sharded - unsharded| = {err:.4f}")
    def t(fn, n=20):
        for _ in range(5): fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize(); s.record()
        for _ in range(n): fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    tn = t(lambda: nccl(x, idm, cd, [ones, ones]))
    tp = t(lambda: p2p(x, idm, cd, [ones, ones]))
    t1 = t(lambda: layer(x, idm, cd, [ones, ones]))
if rank == 0:
    print(f"world {world}: unsharded {t1:.3f} ms, channel-sharded nccl {tn:.3f} ms, p2p push {tp:.3f} ms")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
