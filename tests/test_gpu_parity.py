"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Tolerances (stated per dtype):
  fp32 I/O : rtol 2e-4, atol 2e-5 against the oracle evaluated in float64 (MUFU ex2/lg2 are ~2^-22 accurate
             and the recurrence compounds them);
  bf16 I/O : rtol 1.6e-2, atol 1e-2 against the fp32 oracle on the same bf16 inputs (one output rounding, 2^-8);
  fp16 I/O : rtol 2e-3, atol 2e-3.
Index lists and everything integer are compared with torch.equal.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import LAYER_CASES, VARIANT_CASES, load_golden
from oracle import SS2D_Unit_ref, SS2D_cond_v10_ref, mask_to_index as oracle_mask_to_index, selective_scan_ref

pytestmark = pytest.mark.gpu

TOL = {torch.float32: (2e-4, 2e-5), torch.bfloat16: (1.6e-2, 1e-2), torch.float16: (2e-3, 2e-3)}
LAYER_TOL = {torch.float32: (1e-3, 1e-4), torch.bfloat16: (3e-2, 3e-2), torch.float16: (5e-3, 5e-3)}


def close(got, want, dtype, tol=TOL, what=""):
    rtol, atol = tol[dtype]
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    worst = (err - bound).max().item()
    assert worst <= 0, f"{what}: max abs err {err.max().item():.3e}, worst excess over tol {worst:.3e}"


def scan_inputs(batch, dim, L, G, dtype, seed=0, N=16, structured=False, bias_shift=0.0):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(batch, dim, L, generator=g).to(dtype)
    delta = (torch.randn(batch, dim, L, generator=g) + bias_shift).to(dtype)
    A_log = torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(dim, 1)
    if not structured:
        A_log = A_log + 0.5 * torch.randn(dim, N, generator=g)
    A = -torch.exp(A_log)
    B = torch.randn(batch, G, N, L, generator=g).to(dtype)
    C = torch.randn(batch, G, N, L, generator=g).to(dtype)
    D = 1.0 + 0.2 * torch.randn(dim, generator=g)
    bias = torch.randn(dim, generator=g) - 2.0
    z = torch.randn(batch, dim, L, generator=g).to(dtype)
    return u, delta, A, B, C, D, bias, z


def run_fn(args, dev="cuda", **kw):
    from actalker_b200 import selective_scan_fn
    u, delta, A, B, C, D, bias, z = [t.to(dev) for t in args]
    use_z = kw.pop("use_z", False)
    use_D = kw.pop("use_D", True)
    use_bias = kw.pop("use_bias", True)
    return selective_scan_fn(u, delta, A, B, C, D if use_D else None, z if use_z else None,
                             bias if use_bias else None, **kw)


def run_ref(args, dtype, **kw):
    u, delta, A, B, C, D, bias, z = args
    use_z = kw.pop("use_z", False)
    use_D = kw.pop("use_D", True)
    use_bias = kw.pop("use_bias", True)
    kw.pop("a_kind", None)
    cd = torch.float64 if dtype == torch.float32 else torch.float32
    return selective_scan_ref(u, delta, A, B, C, D if use_D else None, z if use_z else None,
                              bias if use_bias else None, compute_dtype=cd, **kw)


# ------------------------------------------------------------------------------------------ operator seam
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 128, 77, 2), (1, 100, 33, 2), (3, 64, 1, 1), (2, 192, 160, 3), (1, 8, 31, 4)])
def test_selective_scan_fn_matches_oracle(dtype, shape):
    batch, dim, L, G = shape
    args = scan_inputs(batch, dim, L, G, dtype, seed=L)
    got = run_fn(args, delta_softplus=True)
    assert got.dtype == dtype and got.shape == (batch, dim, L)
    close(got, run_ref(args, dtype, delta_softplus=True), dtype, what=f"scan {shape}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("opts", [dict(use_z=True, delta_softplus=True), dict(use_D=False, delta_softplus=True),
                                  dict(use_bias=False, delta_softplus=True),
                                  dict(use_z=True, use_D=False, use_bias=False, delta_softplus=True)])
def test_selective_scan_fn_optional_arguments(dtype, opts):
    args = scan_inputs(2, 96, 70, 2, dtype, seed=5)
    close(run_fn(args, **dict(opts)), run_ref(args, dtype, **dict(opts)), dtype, what=str(opts))


def test_selective_scan_fn_without_softplus_and_last_state():
    args = list(scan_inputs(2, 64, 45, 2, torch.float32, seed=9))
    args[1] = args[1].abs() * 0.1            # raw positive delta, as callers without softplus must provide
    got, last = run_fn(args, delta_softplus=False, return_last_state=True, use_bias=False)
    want, want_last = run_ref(args, torch.float32, delta_softplus=False, return_last_state=True, use_bias=False)
    close(got, want, torch.float32, what="no softplus")
    close(last, want_last, torch.float32, what="last state")
    assert last.dtype == torch.float32 and last.shape == (2, 64, 16)


@pytest.mark.parametrize("L", [2047, 2048, 2049, 5217])
def test_selective_scan_fn_long_sequences(L):
    args = scan_inputs(1, 64, L, 2, torch.float32, seed=L)
    close(run_fn(args, delta_softplus=True), run_ref(args, torch.float32, delta_softplus=True), torch.float32,
          what=f"L={L}")


def test_selective_scan_fn_softplus_threshold_region():
    args = list(scan_inputs(1, 64, 40, 1, torch.float32, seed=2))
    args[1] = torch.linspace(-30.0, 30.0, 40).repeat(1, 64, 1) + 0.25 * args[1]   # sweeps across 20 and far below
    args[6] = torch.zeros(64)
    close(run_fn(args, delta_softplus=True), run_ref(args, torch.float32, delta_softplus=True), torch.float32,
          what="softplus sweep")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_power_path_equals_general_path_and_oracle(dtype):
    from actalker_b200 import a_kind_of, _lib
    args = scan_inputs(2, 128, 300, 2, dtype, seed=11, structured=True)
    assert a_kind_of(args[2].cuda()) == _lib.ACTK_A_POWER
    assert a_kind_of(scan_inputs(1, 64, 4, 1, dtype, seed=1)[2].cuda()) == _lib.ACTK_A_GENERAL
    want = run_ref(args, dtype, delta_softplus=True)
    for kind in ("general", "power", "auto"):
        close(run_fn(args, delta_softplus=True, a_kind=kind), want, dtype, what=kind)


@pytest.mark.parametrize("N", [1, 8, 24, 64])
def test_generic_dstate_kernel(N):
    args = scan_inputs(2, 40, 50, 2, torch.float32, seed=N, N=N)
    close(run_fn(args, delta_softplus=True, use_z=True), run_ref(args, torch.float32, delta_softplus=True, use_z=True),
          torch.float32, what=f"dstate={N}")


def test_strided_views_are_accepted_like_upstream():
    args = list(scan_inputs(2, 64, 50, 2, torch.bfloat16, seed=4))
    big = torch.randn(2, 50, 64).to(torch.bfloat16)
    args[0] = big.permute(0, 2, 1)                       # last stride != 1 -> made contiguous, as mamba-ssm does
    xdbl = torch.randn(2, 2, 40, 50).to(torch.bfloat16)
    args[3], args[4] = xdbl[:, :, 4:20], xdbl[:, :, 20:36]   # split views with a larger group stride
    close(run_fn(args, delta_softplus=True), run_ref(args, torch.bfloat16, delta_softplus=True), torch.bfloat16,
          what="strided")


# ------------------------------------------------------------------------------------------ mask indexing
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_mask_index_bit_exact(dtype):
    from actalker_b200 import mask_to_index
    from actalker_b200.mask import MaskIndexCache
    masks = {"ones": torch.ones(1, 1, 576, 576), "zeros": torch.zeros(1, 1, 576, 576)}
    rect = torch.zeros(1, 1, 576, 576)
    rect[:, :, 300:480, 180:400] = 1.0
    masks["rect"] = rect
    k = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0]) / 16.0
    masks["soft"] = F.conv2d(rect, (k[:, None] * k[None, :])[None, None], padding=2)
    cache = MaskIndexCache()
    for name, m in masks.items():
        m = m.to(dtype)
        for L in (5184, 1296, 324):
            want = oracle_mask_to_index(m, L)
            got = mask_to_index(m.cuda(), L)
            assert torch.equal(got.cpu(), want), (name, L)
            e = cache.get(m.cuda(), L)
            assert e.n_sel == want.numel() and torch.equal(e.idx.cpu().long(), want)
            assert torch.equal(e.selected.cpu().nonzero().view(-1), want)
    misses = cache.misses
    cache.get(m.cuda(), 324)          # different storage -> miss; same tensor -> hit
    mg = m.cuda()
    cache.get(mg, 324); before = cache.misses; cache.get(mg, 324)
    assert cache.misses == before and before > misses


# ------------------------------------------------------------------------------------------ module seam
def build_pair(g, device="cuda"):
    import actalker_b200
    d_model, d_cond, side, _ = g["meta"]
    kw = dict(d_model=d_model, d_cond=d_cond, cond_size=32, dropout=0.1, d_state=16, size=side,
              scan_type="sweep", num_direction=2)
    ours = getattr(actalker_b200, g.get("cls", "SS2D_cond_v10"))(**kw).eval()
    if g["dtype"] != torch.float32:
        ours = ours.to(g["dtype"])
    ours.load_state_dict(g["sd"], strict=True)       # same keys / shapes as the reference (Appendix C)
    ours = ours.to(device)
    for name, p in ours.named_parameters():          # Inference.py:430-433
        if any(s in name for s in ("A_logs", "Ds", "dt_projs_bias")):
            p.data = p.data.float()
    return ours


@pytest.mark.parametrize("case", LAYER_CASES + VARIANT_CASES)
def test_layer_matches_reference_golden(case):
    g = load_golden(case)
    layer = build_pair(g)
    dev = "cuda"
    with torch.no_grad():
        y = layer(g["x"].to(dev), g["id_emb"].to(dev), g["conds"].to(dev), [g["mask0"].to(dev), g["mask1"].to(dev)])
    assert y.dtype == g["dtype"] and y.shape == g["y"].shape
    L = g["x"].shape[1]
    assert torch.equal(layer.mask_cache.get(g["mask0"].to(dev), L).idx64.cpu(), g["idx0"])
    assert torch.equal(layer.mask_cache.get(g["mask1"].to(dev), L).idx64.cpu(), g["idx1"])
    close(y, g["y"], g["dtype"], tol=LAYER_TOL, what=case)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cfg", [(64, 128, 12, 3, "rect"), (320, 1024, 16, 2, "ones"), (96, 64, 10, 2, "split")])
def test_layer_matches_oracle_on_seeded_inputs(dtype, cfg):
    from actalker_b200 import SS2D_cond_v10
    d_model, d_cond, side, batch, mkind = cfg
    torch.manual_seed(72589 + d_model)
    kw = dict(d_model=d_model, d_cond=d_cond, cond_size=32, dropout=0.1, d_state=16, size=side,
              scan_type="sweep", num_direction=2)
    ref = SS2D_cond_v10_ref(**kw).eval()
    with torch.no_grad():
        for unit in (ref.audio_unit, ref.exp_unit):
            unit.A_logs.add_(0.5 * torch.randn_like(unit.A_logs))
            unit.Ds.copy_(1.0 + 0.2 * torch.randn_like(unit.Ds))
    ours = SS2D_cond_v10(**kw).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    if dtype != torch.float32:
        ref, ours = ref.to(dtype), ours.to(dtype)
        for m in (ref, ours):
            for name, p in m.named_parameters():
                if any(s in name for s in ("A_logs", "Ds", "dt_projs_bias")):
                    p.data = p.data.float()
    ours = ours.cuda()
    L, px = side * side, side * 8
    x = torch.randn(batch, L, d_model).to(dtype)
    id_emb = torch.randn(batch, 1, d_cond).to(dtype)
    conds = torch.randn(batch, 33, d_cond).to(dtype)
    ones = torch.ones(1, 1, px, px)
    top = torch.zeros(1, 1, px, px); top[:, :, : px // 2] = 1
    bot = torch.zeros(1, 1, px, px); bot[:, :, px // 2:, px // 4: 3 * px // 4] = 1
    masks = {"ones": [ones, ones], "rect": [bot, ones], "split": [bot, top]}[mkind]
    masks = [m.to(dtype) for m in masks]
    with torch.no_grad():
        want = ref(x.clone(), id_emb, conds, masks)
        got = ours(x.cuda(), id_emb.cuda(), conds.cuda(), [m.cuda() for m in masks])
    close(got, want, dtype, tol=LAYER_TOL, what=f"{cfg} {dtype}")


def test_unit_matches_reference_golden():
    from actalker_b200 import SS2D_Unit
    g = load_golden("unit_f32")
    d_model, L, batch = g["meta"]
    unit = SS2D_Unit(d_model, 64, 32, 16, size=8, scan_type="sweep", num_direction=2).eval()
    unit.load_state_dict(g["sd"], strict=True)
    unit = unit.cuda()
    with torch.no_grad():
        y = unit(g["x"].cuda())
    close(y, g["y"], torch.float32, tol=LAYER_TOL, what="unit")


def test_layer_is_forward_only_and_says_so():
    from actalker_b200 import SS2D_cond_v10
    layer = SS2D_cond_v10(d_model=32, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=8,
                          scan_type="sweep", num_direction=2).cuda()
    ones = torch.ones(1, 1, 64, 64, device="cuda")
    with pytest.raises(NotImplementedError):
        layer(torch.randn(1, 64, 32, device="cuda"), torch.randn(1, 1, 64, device="cuda"),
              torch.randn(1, 33, 64, device="cuda"), [ones, ones])


def test_host_streamed_layer_matches_direct_calls():
    """The host-buffer API (three streams, double-buffered staging) returns what direct device calls return,
    for every submission, including when consecutive submissions reuse a staging slot."""
    from actalker_b200 import SS2D_cond_v10
    from actalker_b200.host_api import HostStreamedLayer
    torch.manual_seed(11)
    layer = SS2D_cond_v10(d_model=64, d_cond=128, cond_size=32, dropout=0.1, d_state=16, size=12,
                          scan_type="sweep", num_direction=2).eval().cuda()
    ones = torch.ones(1, 1, 96, 96, device="cuda")
    work = [(torch.randn(3, 144, 64).pin_memory(), torch.randn(3, 1, 128).pin_memory(),
             torch.randn(3, 33, 128).pin_memory(), torch.empty(3, 144, 64).pin_memory()) for _ in range(5)]
    # inputs staged together in ONE pinned buffer travel as one copy (the bench's e2e path): same results
    for _ in range(3):
        buf = torch.empty(3 * 144 * 64 + 3 * 128 + 3 * 33 * 128).pin_memory()
        px, pi, pc = buf[:27648].view(3, 144, 64), buf[27648:28032].view(3, 1, 128), buf[28032:].view(3, 33, 128)
        for t in (px, pi, pc):
            t.copy_(torch.randn(t.shape))
        assert HostStreamedLayer._packed_base(px, pi, pc) is buf
        work.append((px, pi, pc, torch.empty(3, 144, 64).pin_memory()))
    runner = HostStreamedLayer(layer)
    for x, idm, cd, out in work:
        runner.submit(x, idm, cd, [ones, ones], out)
    runner.drain()
    with torch.no_grad():
        for x, idm, cd, out in work:
            want = layer(x.cuda(), idm.cuda(), cd.cuda(), [ones, ones]).cpu()
            assert torch.equal(out, want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("nseg", [2, 3, 7])
def test_two_level_scan_equals_single_level(dtype, nseg):
    """Chunk-split invariance: cutting every sequence into nseg chunks (summary pass + carry + rescan) must give
    the single-level result up to fp32 re-association, for both directions, ragged tails and partial masks."""
    from actalker_b200 import SS2D_cond_v10, mamba_layer as ml
    torch.manual_seed(5)
    side = 20
    layer = SS2D_cond_v10(d_model=64, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=side,
                          scan_type="sweep", num_direction=2).eval()
    with torch.no_grad():
        layer.exp_unit.A_logs.add_(0.4 * torch.randn_like(layer.exp_unit.A_logs))   # one POWER, one general branch
    if dtype != torch.float32:
        layer = layer.to(dtype)
        for n, p in layer.named_parameters():
            if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
                p.data = p.data.float()
    layer = layer.cuda()
    Bp, L = 2, side * side
    x = torch.randn(Bp, L, 64, device="cuda").to(dtype)
    idm = torch.randn(Bp, 1, 64, device="cuda").to(dtype)
    cd = torch.randn(Bp, 33, 64, device="cuda").to(dtype)
    rect = torch.zeros(1, 1, 160, 160, device="cuda", dtype=dtype)
    rect[:, :, 24:140, 16:120] = 1
    masks = [torch.ones(1, 1, 160, 160, device="cuda", dtype=dtype), rect]
    try:
        with torch.no_grad():
            ml.SCAN_SEGMENTS = 1
            want = layer(x, idm, cd, masks)
            ml.SCAN_SEGMENTS = nseg
            got = layer(x, idm, cd, masks)
    finally:
        ml.SCAN_SEGMENTS = None
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert torch.allclose(got.float(), want.float(), rtol=tol, atol=tol), (got.float() - want.float()).abs().max()


def test_small_batch_long_sequence_picks_two_level_scan_and_matches_oracle():
    """B'=1 with a long flattened sequence (the shape of BASELINE config 5) leaves the GPU empty on the independent
    axes alone; auto_segments must cut time, and the result must still match the CPU oracle."""
    from actalker_b200 import SS2D_cond_v10
    from actalker_b200.mamba_layer import auto_segments
    assert auto_segments(1000, 327) == 1                    # config 2: 1000 CTAs -> single level
    assert auto_segments(40, 32400) == 23                   # config 5 flattened, one sequence
    assert auto_segments(40, 12) == 1                       # too short to cut
    assert auto_segments(500, 327) == 1                     # a rank's half of d_inner: single level is faster
    torch.manual_seed(9)
    side = 48
    kw = dict(d_model=32, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=side, scan_type="sweep",
              num_direction=2)
    ref = SS2D_cond_v10_ref(**kw).eval()
    ours = SS2D_cond_v10(**kw).eval()
    ours.load_state_dict(ref.state_dict())
    ours = ours.cuda()
    L = side * side
    x, idm, cd = torch.randn(1, L, 32), torch.randn(1, 1, 64), torch.randn(1, 33, 64)
    ones = torch.ones(1, 1, 96, 96)
    with torch.no_grad():
        want = ref(x.clone(), idm, cd, [ones, ones])
        got = ours(x.cuda(), idm.cuda(), cd.cuda(), [ones.cuda(), ones.cuda()])
    close(got, want, torch.float32, tol=LAYER_TOL, what="two-level auto")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("chunks", [2, 5])
def test_chained_chunks_are_bit_identical_to_single_level(dtype, chunks):
    """Chain mode hands the fp32 state from chunk to chunk unchanged and performs the same operations in the same
    order, so the result must be bit-identical to the single-level scan (both directions, ragged tiles, masks)."""
    from actalker_b200 import SS2D_cond_v10, mamba_layer as ml
    torch.manual_seed(6)
    side = 24
    layer = SS2D_cond_v10(d_model=96, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=side,
                          scan_type="sweep", num_direction=2).eval()
    with torch.no_grad():
        layer.audio_unit.A_logs.add_(0.4 * torch.randn_like(layer.audio_unit.A_logs))
    if dtype != torch.float32:
        layer = layer.to(dtype)
        for n, p in layer.named_parameters():
            if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
                p.data = p.data.float()
    layer = layer.cuda()
    Bp, L = 3, side * side
    x = torch.randn(Bp, L, 96, device="cuda").to(dtype)
    idm = torch.randn(Bp, 1, 64, device="cuda").to(dtype)
    cd = torch.randn(Bp, 33, 64, device="cuda").to(dtype)
    rect = torch.zeros(1, 1, 192, 192, device="cuda", dtype=dtype)
    rect[:, :, 24:170, 16:150] = 1
    masks = [torch.ones(1, 1, 192, 192, device="cuda", dtype=dtype), rect]
    try:
        with torch.no_grad():
            ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = 1, 0
            want = layer(x, idm, cd, masks)
            ml.SCAN_CHAIN = chunks
            got = [layer(x, idm, cd, masks) for _ in range(3)]
    finally:
        ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = None, None
    for g in got:
        assert torch.equal(g, want)


@pytest.mark.parametrize("cfg", [(64, 12, 2, torch.float32), (320, 18, 4, torch.bfloat16), (320, 72, 25, torch.bfloat16)])
def test_layer_is_cuda_graph_capturable(cfg):
    """After the caches are warm a layer call has static shapes and no host sync: capture it (tensor maps, chain-mode
    memset and all) and replay with new input contents; results must equal eager calls bit for bit."""
    from actalker_b200 import SS2D_cond_v10
    from actalker_b200.graphed import GraphedLayer
    d_model, side, Bp, dtype = cfg
    torch.manual_seed(13)
    layer = SS2D_cond_v10(d_model=d_model, d_cond=128, cond_size=32, dropout=0.1, d_state=16, size=side,
                          scan_type="sweep", num_direction=2).eval()
    if dtype != torch.float32:
        layer = layer.to(dtype)
        for n, p in layer.named_parameters():
            if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
                p.data = p.data.float()
    layer = layer.cuda()
    L = side * side
    mk = lambda: (torch.randn(Bp, L, d_model, device="cuda").to(dtype), torch.randn(Bp, 1, 128, device="cuda").to(dtype),
                  torch.randn(Bp, 33, 128, device="cuda").to(dtype))
    rect = torch.zeros(1, 1, side * 8, side * 8, device="cuda", dtype=dtype)
    rect[:, :, side: 7 * side, 2 * side: 6 * side] = 1
    masks = [torch.ones_like(rect), rect]
    g = GraphedLayer(layer, *mk(), masks)
    for _ in range(3):
        x, idm, cd = mk()
        got = g(x, idm, cd).clone()
        with torch.no_grad():
            want = layer(x, idm, cd, masks)
        assert torch.equal(got, want)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mode", ["single", "chain", "two_level"])
@pytest.mark.parametrize("cfg", [(64, 20, 3), (320, 24, 2), (640, 12, 2), (1280, 10, 1), (96, 72, 25)])
def test_fused_dt_proj_equals_gemm_route(dtype, cfg, mode):
    """with 16-bit activations the scan kernel can compute dt_proj itself per 16-token tile
    (tcgen05.mma into tensor memory, fp32 accumulate, one rounding to the activation dtype), in all three launch
    shapes (single level, chained chunks, two-level).  Same contraction and rounding point as the cuBLAS GEMM
    that otherwise writes the delta tensor, so the two routes may only differ where the fp32 sums round differently:
    require agreement within one 16-bit rounding step of the layer output, on ragged masks, the id/cond tail, a
    partial channel block (d_model 96 -> D = 192) and every rank slab count (dt_rank 4 / 20 / 40 / 80)."""
    from actalker_b200 import SS2D_cond_v10, mamba_layer as ml
    d_model, side, Bp = cfg
    torch.manual_seed(21)
    layer = SS2D_cond_v10(d_model=d_model, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=side,
                          scan_type="sweep", num_direction=2).eval()
    with torch.no_grad():
        layer.exp_unit.A_logs.add_(0.3 * torch.randn_like(layer.exp_unit.A_logs))
    layer = layer.to(dtype)
    for n, p in layer.named_parameters():
        if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
            p.data = p.data.float()
    layer = layer.cuda()
    assert layer.audio_unit.derived()["fusable"]
    L = side * side
    x = torch.randn(Bp, L, d_model, device="cuda").to(dtype)
    idm = torch.randn(Bp, 1, 64, device="cuda").to(dtype)
    cd = torch.randn(Bp, 33, 64, device="cuda").to(dtype)
    rect = torch.zeros(1, 1, side * 8, side * 8, device="cuda", dtype=dtype)
    rect[:, :, side: 7 * side, 2 * side: 6 * side + 3] = 1
    masks = [torch.ones_like(rect), rect]
    was = ml.FUSE_DT_PROJ
    try:
        ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = {"single": (1, 0), "chain": (1, 3), "two_level": (3, 0)}[mode]
        with torch.no_grad():
            ml.FUSE_DT_PROJ = False
            want = layer(x, idm, cd, masks)
            ml.FUSE_DT_PROJ = True
            got = layer(x, idm, cd, masks)
    finally:
        ml.FUSE_DT_PROJ, ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = was, None, None
    assert torch.isfinite(got.float()).all()
    step = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    err = (got.float() - want.float()).abs()
    assert (err <= step * (1.0 + want.float().abs())).all(), err.max()
    assert (got == want).float().mean() > 0.97    # almost every output element is bit-identical


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cfg", [(18, 3, "v10"), (24, 2, "v10"), (13, 5, "v10"), (16, 2, "v8"), (72, 25, "v10")])
def test_fused_ln_out_proj_equals_unfused_route(dtype, cfg):
    """merge + LayerNorm + out_proj as one tcgen05 kernel (D = 640) against the merge kernel + cuBLAS
    route.  Both round the normalised tensor to the activation dtype and accumulate the projection in fp32, so they
    may differ only by fp32 summation order before the final rounding: agreement within one rounding step of the
    output, on partial row tiles (B'L % 128 != 0), rectangle masks (pass-through rows) and the v8 row weights."""
    from actalker_b200 import SS2D_cond_v10, SS2D_cond_v8, mamba_layer as ml
    side, Bp, kind = cfg
    torch.manual_seed(33)
    cls = SS2D_cond_v8 if kind == "v8" else SS2D_cond_v10
    layer = cls(d_model=320, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=side, scan_type="sweep",
                num_direction=2).eval()
    with torch.no_grad():
        layer.out_norm.weight.add_(0.2 * torch.randn_like(layer.out_norm.weight))
        layer.out_norm.bias.add_(0.2 * torch.randn_like(layer.out_norm.bias))
    layer = layer.to(dtype)
    for n, p in layer.named_parameters():
        if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
            p.data = p.data.float()
    layer = layer.cuda()
    L = side * side
    x = torch.randn(Bp, L, 320, device="cuda").to(dtype)
    idm = torch.randn(Bp, 1, 64, device="cuda").to(dtype)
    cd = torch.randn(Bp, 33, 64, device="cuda").to(dtype)
    rect = torch.zeros(1, 1, side * 8, side * 8, device="cuda", dtype=dtype)
    rect[:, :, side: 7 * side, 2 * side: 6 * side + 3] = 1
    masks = [torch.ones_like(rect), rect]
    was = ml.FUSE_LN_OUT_PROJ
    try:
        with torch.no_grad():
            ml.FUSE_LN_OUT_PROJ = False
            want = layer(x, idm, cd, masks)
            ml.FUSE_LN_OUT_PROJ = True
            got = layer(x, idm, cd, masks)
    finally:
        ml.FUSE_LN_OUT_PROJ = was
    assert got.shape == want.shape and torch.isfinite(got.float()).all()
    step = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    err = (got.float() - want.float()).abs()
    assert (err <= step * (1.0 + want.float().abs())).all(), err.max()
    assert (got == want).float().mean() > 0.9


def _random_case(i):
    """Seeded random layer configuration: size, batch, dtype, token-aligned rectangle masks (possibly empty or full) and
    a forced launch shape.  Rectangles are aligned to the 8-pixel token grid, where the bicubic downsample is exactly
    0 / 1 on every device, so the index lists of the CPU oracle and the GPU layer cannot differ by rounding."""
    rng = np.random.RandomState(1000 + i)
    d_model = int(rng.choice([32, 64, 96, 160, 320]))
    side = int(rng.randint(3, 21))
    Bp = int(rng.randint(1, 4))
    dtype = [torch.float32, torch.bfloat16, torch.float16][i % 3]
    masks = []
    for _ in range(2):
        kind = rng.randint(0, 5)
        m = torch.zeros(1, 1, side * 8, side * 8)
        if kind == 0:
            m[:] = 1
        elif kind == 1:
            pass                                           # nothing selected: the branch is the in_proj pass-through
        else:
            r0, c0 = rng.randint(0, side), rng.randint(0, side)
            r1, c1 = rng.randint(r0 + 1, side + 1), rng.randint(c0 + 1, side + 1)
            m[:, :, 8 * r0:8 * r1, 8 * c0:8 * c1] = 1
        masks.append(m)
    shape = [(None, None), (1, 0), (1, 2), (1, 3), (2, 0), (3, 0)][rng.randint(0, 6)]
    return d_model, side, Bp, dtype, masks, shape


@pytest.mark.parametrize("i", range(36))
def test_layer_random_sweep_matches_oracle(i):
    """Randomised end-to-end parity of the drop-in layer against the CPU oracle over sizes, batches, dtypes, mask
    geometries (including empty selections and single-row rectangles) and all three launch shapes of the scan."""
    from actalker_b200 import SS2D_cond_v10, mamba_layer as ml
    d_model, side, Bp, dtype, masks, (seg, chain) = _random_case(i)
    torch.manual_seed(5000 + i)
    kw = dict(d_model=d_model, d_cond=48, cond_size=32, dropout=0.1, d_state=16, size=side, scan_type="sweep",
              num_direction=2)
    ref = SS2D_cond_v10_ref(**kw).eval()
    with torch.no_grad():
        ref.exp_unit.A_logs.add_(0.4 * torch.randn_like(ref.exp_unit.A_logs))
        ref.audio_unit.Ds.copy_(1.0 + 0.3 * torch.randn_like(ref.audio_unit.Ds))
        ref.out_norm.weight.add_(0.1 * torch.randn_like(ref.out_norm.weight))
    ours = SS2D_cond_v10(**kw).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    if dtype != torch.float32:
        ref, ours = ref.to(dtype), ours.to(dtype)
        for m in (ref, ours):
            for name, p in m.named_parameters():
                if any(s in name for s in ("A_logs", "Ds", "dt_projs_bias")):
                    p.data = p.data.float()
    ours = ours.cuda()
    L = side * side
    x = torch.randn(Bp, L, d_model).to(dtype)
    id_emb = torch.randn(Bp, 1, 48).to(dtype)
    conds = torch.randn(Bp, 33, 48).to(dtype)
    masks = [m.to(dtype) for m in masks]
    try:
        ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = seg, chain
        with torch.no_grad():
            want = ref(x.clone(), id_emb, conds, masks)
            got = ours(x.cuda(), id_emb.cuda(), conds.cuda(), [m.cuda() for m in masks])
    finally:
        ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = None, None
    close(got, want, dtype, tol=LAYER_TOL, what=f"random case {i}: d_model {d_model} side {side} B' {Bp} {dtype} shape {(seg, chain)}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("use_z", [False, True])
def test_operator_matches_the_upstream_derived_cuda_kernel_in_vllm(dtype, use_z):
    """Independent pin of the operator contract: vLLM ships a CUDA selective scan derived from mamba-ssm's own kernel
    (csrc/mamba/mamba_ssm/selective_scan_fwd.cu, bound as vllm._custom_ops.selective_scan_fwd) — the closest thing to the
    reference's `selective_scan_fn` (mamba-ssm 1.2.0.post1 is not installable here) that exists on the box.  Same
    inputs through both kernels: grouped B/C, D skip, delta bias, softplus, optional SiLU(z) gate, last state."""
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn as vllm_scan
    except Exception as e:   # noqa: BLE001
        pytest.skip(f"vllm's mamba kernel is not importable here: {type(e).__name__}")
    from actalker_b200 import selective_scan_fn
    torch.manual_seed(11)
    batch, dim, L, N = 3, 256, 517, 16
    dev = "cuda"
    u = torch.randn(batch, dim, L, device=dev).to(dtype)
    delta = (0.5 * torch.randn(batch, dim, L, device=dev)).to(dtype)
    A = -torch.exp(torch.randn(dim, N, device=dev) * 0.5)
    Bm = torch.randn(batch, 1, N, L, device=dev).to(dtype)
    Cm = torch.randn(batch, 1, N, L, device=dev).to(dtype)
    D = torch.randn(dim, device=dev)
    bias = torch.randn(dim, device=dev) - 2.0
    z = torch.randn(batch, dim, L, device=dev).to(dtype) if use_z else None
    ours, last = selective_scan_fn(u, delta, A, Bm, Cm, D, z, bias, True, return_last_state=True)
    state = torch.zeros(batch, dim, N, device=dev, dtype=dtype)
    try:
        theirs = vllm_scan(u.clone(), state, delta.clone(), A, Bm.clone(), Cm.clone(), D, None if z is None else z.clone(),
                           bias, True, has_initial_state=torch.zeros(batch, dtype=torch.bool, device=dev))
    except Exception as e:   # noqa: BLE001 - the op may be compiled out of this vllm build
        pytest.skip(f"vllm selective_scan_fwd not runnable here: {type(e).__name__}: {str(e)[:120]}")
    close(ours, theirs, dtype, what="vs vllm mamba kernel")
    rt, at = TOL[dtype]
    assert torch.allclose(last.float(), state.float(), rtol=max(rt, 2e-3), atol=max(at, 2e-3)), \
        (last.float() - state.float()).abs().max()
