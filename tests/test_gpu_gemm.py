"""The tensor-core projection kernel (C-ABI actk_gemm_tn_fwd) against a plain fp32 torch reference of the same op.

C = round(A @ W^T) with fp32 accumulation and one rounding — the contract of the reference's nn.Linear / einsum calls
(src/models/base/mamba_layer.py:1960-1961, :1521-1523, :1985).  The fp32 reference is evaluated in float64 on the same
16-bit inputs and rounded once, so the two may differ only where fp32 summation order moves a value across a rounding
boundary: |err| <= one 16-bit rounding step of the result (tolerance written below).
"""
import ctypes as ct

import pytest
import torch

pytestmark = pytest.mark.gpu

STEP = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11}   # half an ulp relative to the value, doubled below


def _ref(a, w, dtype, silu=False):
    y = (a.double() @ w.double().t()).float().to(dtype)
    if silu:
        y = torch.nn.functional.silu(y.float()).to(dtype)
    return y


def _check(got, want, dtype, what, ulps=1):
    assert got.dtype == dtype and got.shape == want.shape, what
    assert torch.isfinite(got.float()).all(), f"{what}: non-finite"
    err = (got.float() - want.float()).abs()
    tol = ulps * 2 * STEP[dtype] * want.float().abs() + 1e-3 * STEP[dtype] / 2.0 ** -11 + 1e-6
    bad = (err > tol)
    assert not bad.any(), f"{what}: {int(bad.sum())} elements off, max err {err.max().item():.3e}"
    assert (got == want).float().mean() > 0.95, f"{what}: only {(got == want).float().mean().item():.3f} bit-identical"


SHAPES = [
    (129600 // 25 * 2, 640, 320),     # in_proj rows of two frames
    (777, 128, 640),                  # x_proj, ragged M
    (1000, 1280, 64),                 # dt_proj: one K slab
    (300, 2560, 96),                  # dt_proj at d_model 640: K = 2*48, partial second slab
    (513, 320, 640),                  # out_proj: two column tiles of 160
    (1, 640, 1024),                   # a single condition token
    (33, 192, 48),                    # test-sized layer: K < one slab, N = 192
    (260, 72, 40),                    # N not a multiple of 32: last column chunk clipped
    (128, 16, 8),                     # smallest legal row pitch
]


@pytest.fixture(params=["auto", "pair", "rows256"])
def tile_rows(request, monkeypatch):
    """All launch forms of the kernel on every shape: the default (one CTA per 128-row tile, two accumulator buffers), and the
    two opt-in forms that were built to share W slabs between more rows and measured slower — CTA pairs on 256-row tiles
    (ACTK_GEMM_PAIR=1, cta_group::2: clusters of two, each CTA loads its 128 rows of A and half of every W slab, the leader
    issues M = 256 MMAs) and 256-row tiles in one CTA (ACTK_GEMM_MH=2: every W slab feeds two M = 128 MMAs)."""
    monkeypatch.delenv("ACTK_GEMM_PAIR", raising=False)
    monkeypatch.delenv("ACTK_GEMM_MH", raising=False)
    if request.param == "pair":
        monkeypatch.setenv("ACTK_GEMM_PAIR", "1")
    elif request.param == "rows256":
        monkeypatch.setenv("ACTK_GEMM_MH", "2")
    return request.param


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_matches_fp32_reference(shape, dtype, tile_rows):
    from actalker_b200 import gemm
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dtype).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dtype).cuda()
    out = torch.full((M, N), float("nan"), dtype=dtype, device="cuda")
    gemm.run([gemm.Problem(a, w, out)])
    _check(out.cpu(), _ref(a.cpu(), w.cpu(), dtype), dtype, f"gemm {shape}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_silu_epilogue_planes_and_strided_operands(dtype, tile_rows):
    from actalker_b200 import gemm
    g = torch.Generator().manual_seed(3)
    # (1) SiLU epilogue: round(silu(round(a @ w^T)))
    a = torch.randn(700, 128, generator=g).to(dtype).cuda()
    w = (torch.randn(320, 128, generator=g) / 11.0).to(dtype).cuda()
    out = torch.full((700, 320), float("nan"), dtype=dtype, device="cuda")
    gemm.run([gemm.Problem(a, w, out)], silu=True)
    # two roundings: a product that lands one ulp apart moves the SiLU of it by up to two (slope ~1.1 above x = 2)
    _check(out.cpu(), _ref(a.cpu(), w.cpu(), dtype, silu=True), dtype, "silu", ulps=3)
    # (2) two output planes from one stacked weight (in_proj1 | in_proj2)
    w2 = (torch.randn(2 * 192, 128, generator=g) / 11.0).to(dtype).cuda()
    planes = torch.full((2, 700, 192), float("nan"), dtype=dtype, device="cuda")
    gemm.run([gemm.Problem(a, w2, planes, planes=2)])
    want = _ref(a.cpu(), w2.cpu(), dtype)
    _check(planes[0].cpu(), want[:, :192], dtype, "plane 0")
    _check(planes[1].cpu(), want[:, 192:], dtype, "plane 1")
    # (2b) planes of 640 columns (in_proj at d_model 320): 256-column tiles, the third one straddles the two planes and
    # its 64-column chunks go to different tensors
    w3 = (torch.randn(2 * 640, 128, generator=g) / 11.0).to(dtype).cuda()
    planes = torch.full((2, 700, 640), float("nan"), dtype=dtype, device="cuda")
    gemm.run([gemm.Problem(a, w3, planes, planes=2)])
    want = _ref(a.cpu(), w3.cpu(), dtype)
    _check(planes[0].cpu(), want[:, :640], dtype, "plane 0 of 2 x 640")
    _check(planes[1].cpu(), want[:, 640:], dtype, "plane 1 of 2 x 640")
    # (3) strided A (the dt columns of x_dbl read in place) and strided C (one token slot of the tail buffer)
    xdbl = torch.randn(500, 128, generator=g).to(dtype).cuda()
    wd = (torch.randn(256, 64, generator=g) / 8.0).to(dtype).cuda()
    buf = torch.full((500, 3, 256), float("nan"), dtype=dtype, device="cuda")
    gemm.run([gemm.Problem(xdbl[:, 64:], wd, buf[:, 1, :])])
    _check(buf[:, 1, :].cpu(), _ref(xdbl[:, 64:].cpu(), wd.cpu(), dtype), dtype, "strided")
    assert torch.isnan(buf[:, 0, :].float()).all() and torch.isnan(buf[:, 2, :].float()).all()   # neighbours untouched


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_gemm_grouped_launch_and_many_tiles_per_cta(dtype, tile_rows):
    """Four problems of different sizes in ONE launch (the x_proj / dt_proj launches of the layer), with more tiles than
    SMs so that every CTA walks the ring and both accumulator buffers many times."""
    from actalker_b200 import gemm
    g = torch.Generator().manual_seed(5)
    dims = [(40000, 128, 640), (70, 128, 640), (30011, 224, 320), (9, 160, 1280)]
    probs, wants = [], []
    for M, N, K in dims:
        a = torch.randn(M, K, generator=g).to(dtype).cuda()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dtype).cuda()
        out = torch.full((M, N), float("nan"), dtype=dtype, device="cuda")
        probs.append(gemm.Problem(a, w, out))
        wants.append((a.float() @ w.float().t()).to(dtype))        # fp32 on the GPU: tolerance covers the order
    gemm.run(probs)
    for p, want, d in zip(probs, wants, dims):
        _check(p.out.cpu(), want.cpu(), dtype, f"grouped {d}")


def test_gemm_rejects_bad_arguments_loudly():
    from actalker_b200 import _lib, gemm
    lib = _lib.load()
    a = torch.randn(64, 36, device="cuda").to(torch.bfloat16)          # row pitch 72 bytes: not a multiple of 16
    w = torch.randn(32, 36, device="cuda").to(torch.bfloat16)
    out = torch.empty(64, 32, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="16 bytes"):
        gemm.run([gemm.Problem(a, w, out)])
    arr = (_lib.GemmProblem * 1)()
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_F32, None) != 0            # fp32 is not this kernel's route
    assert lib.actk_gemm_tn_fwd(arr, 5, _lib.ACTK_BF16, None) != 0           # more problems than a launch holds
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) != 0           # NULL pointers
    assert b"NULL" in lib.actk_last_error()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_layer_tensor_core_route_equals_torch_gemm_route(dtype):
    """The whole layer with this repo's projections against the same layer with torch's GEMMs (ACTK_TC_GEMM=0): both
    accumulate in fp32 and round at the same points, so outputs agree within one rounding step, on ragged masks, the
    tail tokens and a width whose column tiles are partial (d_model 96 -> D = 192)."""
    from actalker_b200 import SS2D_cond_v10, mamba_layer as ml
    for d_model, side, Bp in [(96, 20, 3), (320, 24, 2)]:
        torch.manual_seed(17)
        layer = SS2D_cond_v10(d_model=d_model, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=side,
                              scan_type="sweep", num_direction=2).eval()
        layer = layer.to(dtype)
        for n, p in layer.named_parameters():
            if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
                p.data = p.data.float()
        layer = layer.cuda()
        L = side * side
        x = torch.randn(Bp, L, d_model, device="cuda").to(dtype)
        idm = torch.randn(Bp, 1, 64, device="cuda").to(dtype)
        cd = torch.randn(Bp, 33, 64, device="cuda").to(dtype)
        rect = torch.zeros(1, 1, side * 8, side * 8, device="cuda", dtype=dtype)
        rect[:, :, side: 7 * side, 2 * side: 6 * side] = 1
        masks = [torch.ones_like(rect), rect]
        was = ml.TC_GEMM
        try:
            with torch.no_grad():
                ml.TC_GEMM = False
                want = layer(x, idm, cd, masks)
                ml.TC_GEMM = True
                got = layer(x, idm, cd, masks)
        finally:
            ml.TC_GEMM = was
        assert torch.isfinite(got.float()).all()
        # projections feed a recurrence and a LayerNorm: a one-ulp input difference moves outputs by a few ulps
        step = 2.0 ** -6 if dtype == torch.bfloat16 else 2.0 ** -9
        err = (got.float() - want.float()).abs()
        assert (err <= step * (1.0 + want.float().abs())).all(), (d_model, err.max().item())


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_gemm_fused_all_gather_stores_every_tile_to_every_target(dtype):
    """The fused GEMM + all-gather epilogue on ONE GPU: eight 'rank' buffers that all live here.  N = 320 takes 192-column
    tiles (store maps are 64 columns wide), so every second tile ends in a chunk that lies outside the output — the case
    that once let a staging tile be overwritten while the last-issued stores were still reading it."""
    from actalker_b200 import gemm
    g = torch.Generator().manual_seed(9)
    M, N, K = 40000, 320, 640
    a = torch.randn(M, K, generator=g).to(dtype).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dtype).cuda()
    want = torch.empty(M, N, dtype=dtype, device="cuda")
    gemm.run([gemm.Problem(a, w, want)])
    bufs = [torch.full((3 + M, N), float("nan"), dtype=dtype, device="cuda") for _ in range(8)]
    for _ in range(3):
        gemm.run([gemm.Problem(a, w, None, peers=[b[3:].data_ptr() for b in bufs])])
    torch.cuda.synchronize()
    for i, b in enumerate(bufs):
        assert torch.equal(b[3:], want), f"target {i}: {(b[3:] != want).sum().item()} elements differ"
        assert torch.isnan(b[:3].float()).all()
    _check(want.cpu(), _ref(a.cpu(), w.cpu(), dtype), dtype, "fused all-gather reference")
