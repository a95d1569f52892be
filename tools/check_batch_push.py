"""torchrun --nproc-per-node N tools/check_batch_push.py — ONE call split batch-first over the ranks: the fused
out_proj + all-gather (TMA stores into every rank's gather buffer over NVLink peer memory) against the NCCL all-gather
route and against the one-GPU call, with inputs that change from call to call (stale buffers would show), and a per-source
-rank report of any mismatch (which rank's block, how many elements, zero / stale / other)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tools.bench_configs import make_layer
from actalker_b200.sharded import BatchShardedCall

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dtype = torch.bfloat16
layer = make_layer(320)
L = 5184
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
ok = True
lines = []
with torch.no_grad():
    for Bp in (100, 25):
        g = torch.Generator(device="cuda").manual_seed(5)
        x0 = torch.randn(Bp, L, 320, device="cuda", generator=g).to(dtype)
        idm = torch.randn(Bp, 1, 1024, device="cuda", generator=g).to(dtype)
        cd = torch.randn(Bp, 33, 1024, device="cuda", generator=g).to(dtype)
        nccl, push = BatchShardedCall(layer, gather="nccl"), BatchShardedCall(layer, gather="p2p")
        bounds = push.plan(Bp).all_bounds()
        prev = prev2 = None
        for it in range(5):
            x = (x0 * (1.0 + 0.25 * it)).to(dtype)
            a = nccl(x, idm, cd, [ones, ones]).clone()
            view = push(x, idm, cd, [ones, ones])
            b = view.clone()
            torch.cuda.synchronize()
            dist.barrier()
            import time
            time.sleep(0.05)
            again = view.clone()                       # the same buffer read once more, long after every rank has finished
            torch.cuda.synchronize()
            if not torch.equal(a, b):
                lines.append(f"rank {rank} B'={Bp} call {it}: first read wrong; second read (after barrier + 50 ms) "
                             f"{'RIGHT: the writes landed late' if torch.equal(a, again) else 'still wrong: ' + str(int((a != again).sum())) + ' elements'}"
                             + (f"; wrong elements equal to the value of two calls ago: {int((b[a != b] == prev2[a != b]).sum())} of {int((a != b).sum())}" if prev2 is not None else ""))
                ok = False
                for r, (lo, hi) in enumerate(bounds):
                    if hi > lo and not torch.equal(a[lo:hi], b[lo:hi]):
                        bad = (a[lo:hi] != b[lo:hi])
                        rows = bad.view(-1, 320).any(dim=1).nonzero().view(-1)
                        zero = int((b[lo:hi][bad] == 0).sum())
                        stale = int((b[lo:hi][bad] == prev[lo:hi][bad]).sum()) if prev is not None else -1
                        lines.append(f"rank {rank} B'={Bp} call {it}: block of rank {r}: {int(bad.sum())} elements differ in "
                                     f"{rows.numel()} rows (first {int(rows[0])}, last {int(rows[-1])} of {(hi - lo) * L}), "
                                     f"{zero} are zero, {stale} equal the previous call's value")
            prev2, prev = prev, a
        if rank == 0:
            want = layer(x, idm, cd, [ones, ones])
            lines.append(f"B'={Bp}: nccl route == one GPU: {torch.equal(a, want)}  (max diff {(a.float() - want.float()).abs().max().item():.4g})")
        push._peer.close()
for r in range(world):
    dist.barrier()
    if r == rank and lines:
        print("\n".join(lines), flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: fused push == nccl route on every rank and call: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
