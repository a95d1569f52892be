"""Multi-GPU path: the partition plan and the all-gather plumbing on CPU with world_size-2 gloo, and the
channel-sliced compute on one GPU with the ranks emulated one after another (B200_PROFILING.md: never co-run
waiting kernels of several ranks on one GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from actalker_b200.sharded import ShardPlan, all_gather_slices


def test_shard_plan_bounds():
    p = ShardPlan("channel", 8, 640, 8)
    assert p.all_bounds() == [(80 * r, 80 * r + 80) for r in range(8)]
    assert ShardPlan("channel", 2, 2560, 8).bounds(1) == (1280, 2560)
    with pytest.raises(ValueError):
        ShardPlan("channel", 3, 640, 8)            # 640 / 3 is not 16-byte granular
    b = ShardPlan("batch", 8, 25)                   # 25 frames over 8 GPUs: 4 3 3 3 3 3 3 3
    sizes = [hi - lo for lo, hi in b.all_bounds()]
    assert sizes == [4, 3, 3, 3, 3, 3, 3, 3] and b.bounds(0)[0] == 0 and b.bounds(7)[1] == 25
    assert all(b.bounds(r)[1] == b.bounds(r + 1)[0] for r in range(7))
    assert ShardPlan("batch", 4, 100).all_bounds() == [(25 * r, 25 * r + 25) for r in range(4)]
    blk = ShardPlan("block", 8, 100)                # equal blocks of ceil(100/8) = 13 frames, the last one shorter
    assert [hi - lo for lo, hi in blk.all_bounds()] == [13] * 7 + [9] and blk.bounds(7) == (91, 100)
    assert [hi - lo for lo, hi in ShardPlan("block", 8, 25).all_bounds()] == [4, 4, 4, 4, 4, 4, 1, 0]
    with pytest.raises(ValueError):
        ShardPlan("rows", 2, 8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                   # same full tensor on every rank
        Bp, L, D = 3, 10, 32
        full = torch.randn(Bp, L, D)
        plan = ShardPlan("channel", world, D, 8)
        lo, hi = plan.bounds(rank)
        gathered = all_gather_slices(full[..., lo:hi].contiguous())        # (P, Bp, L, Ds)
        assert gathered.shape == (world, Bp, L, (hi - lo))
        for p in range(world):
            plo, phi = plan.bounds(p)
            assert torch.equal(gathered[p], full[..., plo:phi])
        # the gathered layout, read slice-major, is the full tensor: what actk_gathered_layernorm_fwd normalises
        rebuilt = gathered.permute(1, 2, 0, 3).reshape(Bp, L, D)
        assert torch.equal(rebuilt, full)
        want = torch.nn.functional.layer_norm(full, (D,))
        assert torch.allclose(torch.nn.functional.layer_norm(rebuilt, (D,)), want)
        # batch mode: ranks own disjoint frame ranges that tile B'
        bp = ShardPlan("batch", world, 7)
        mine = torch.zeros(7)
        mine[slice(*bp.bounds(rank))] = 1
        dist.all_reduce(mine)
        assert torch.equal(mine, torch.ones(7))
        # ONE call split batch-first (BatchShardedCall): block plan, in-place all-gather into the padded buffer, tiled
        # variant; a stand-in layer that treats frames independently, as the real one does
        from actalker_b200.sharded import BatchShardedCall

        class Fake(torch.nn.Module):
            def forward(self, x, id_emb, conds, masks):
                return x * 2.0 + id_emb.mean(dim=(1, 2), keepdim=True) + conds.sum(dim=(1, 2), keepdim=True) * masks[0].mean()

        fake = Fake()
        for Bp in (7, 3, 1):                                   # 4+3 frames, 2+1, 1+0 (an idle rank)
            x, idm, cd = torch.randn(Bp, 5, 8), torch.randn(Bp, 1, 4), torch.randn(Bp, 3, 4)
            masks = [torch.ones(1, 1, 2, 2), torch.ones(1, 1, 2, 2)]
            want = fake(x, idm, cd, masks)
            for tiles in (1, 2, 3):
                call = BatchShardedCall(fake, tiles=tiles)
                assert call.plan(Bp).bounds(rank) == (min(rank * -(-Bp // world), Bp), min((rank + 1) * -(-Bp // world), Bp))
                got = call(x, idm, cd, masks)
                assert got.shape == want.shape and torch.equal(got, want), (Bp, tiles)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_all_gather_plumbing_with_gloo_world_size_2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("d_model,world", [(64, 2), (320, 8), (96, 4)])
def test_channel_sharded_layer_equals_unsharded(dtype, d_model, world):
    """Ranks emulated sequentially on one GPU: every rank's slice scan + merge, stacked the way the all-gather
    would deliver them, normalised by the gathered-LayerNorm kernel, must reproduce the unsharded layer."""
    from actalker_b200 import SS2D_cond_v10
    from actalker_b200.sharded import ShardedSS2DCondV10
    torch.manual_seed(7)
    side = 10
    layer = SS2D_cond_v10(d_model=d_model, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=side,
                          scan_type="sweep", num_direction=2).eval()
    with torch.no_grad():
        for u in (layer.audio_unit, layer.exp_unit):
            u.A_logs.add_(0.3 * torch.randn_like(u.A_logs))
    if dtype != torch.float32:
        layer = layer.to(dtype)
        for n, p in layer.named_parameters():
            if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
                p.data = p.data.float()
    layer = layer.cuda()
    Bp, L = 3, side * side
    x = torch.randn(Bp, L, d_model, device="cuda").to(dtype)
    idm = torch.randn(Bp, 1, 64, device="cuda").to(dtype)
    cd = torch.randn(Bp, 33, 64, device="cuda").to(dtype)
    rect = torch.zeros(1, 1, 80, 80, device="cuda", dtype=dtype)
    rect[:, :, 16:64, 8:72] = 1
    masks = [torch.ones(1, 1, 80, 80, device="cuda", dtype=dtype), rect]
    wrap = ShardedSS2DCondV10(layer, mode="channel")
    plan = ShardPlan("channel", world, layer.d_inner, 8)
    with torch.no_grad():
        want = layer(x, idm, cd, masks)
        proj = layer.project_inputs(x, idm, cd, masks)
        parts = [layer.scan_core(*proj, ch_slice=plan.bounds(r)) for r in range(world)]
        gathered = torch.stack(parts, dim=0)
        got = layer.out_proj(wrap.gathered_layernorm(gathered))
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert torch.allclose(got.float(), want.float(), rtol=tol, atol=tol), (got.float() - want.float()).abs().max()


@pytest.mark.gpu
def test_peer_memory_gather_equals_nccl_route_on_two_gpus():
    """The fused push all-gather (merge kernel storing into every rank's gather buffer over NVLink peer memory) must be
    bit-identical to the NCCL all-gather route and to the unsharded layer.  Needs two visible GPUs (skipped on the
    single-GPU test box; tools/check_p2p.py is the same check for torchrun, recorded under profiles/)."""
    import subprocess
    import sys as _sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "check_p2p.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "p2p == nccl: True" in r.stdout
