"""The lean scan kernel (csrc/masked_scan_lean.cu) against the general one and against the oracle.

Under all-ones masks with 16-bit activations — what the shipped pipeline feeds SS2D_cond_v10 (Inference.py:545-546,
mamba_layer.py:1955-1986) — `actk_masked_scan_fwd` can take the lean kernel (ACTK_LEAN_SCAN=1 / mamba_layer.LEAN_SCAN):
warp-autonomous tiles, fp32 B|C written by the x_proj launch.  It runs the same ChannelScan arithmetic in the same order,
so the layer output must be BIT-IDENTICAL to the general kernel's in every launch shape, and it is checked against the
oracle as well.  (Opt-in: measured no faster than the general kernel)
"""
import pytest
import torch

from test_gpu_parity import LAYER_TOL, close
from test_gpu_unet_widths import _case, _ours

pytestmark = pytest.mark.gpu
LEAN_DEFAULT = False   # mamba_layer.LEAN_SCAN without ACTK_LEAN_SCAN=1 in the environment


def _run(ours, x, id_emb, conds, masks, lean, chain):
    from actalker_b200 import mamba_layer as ml
    try:
        ml.LEAN_SCAN, ml.SCAN_CHAIN, ml.SCAN_SEGMENTS = lean, chain, 1
        with torch.no_grad():
            return ours(x.cuda(), id_emb.cuda(), conds.cuda(), [m.cuda() for m in masks])
    finally:
        ml.LEAN_SCAN, ml.SCAN_CHAIN, ml.SCAN_SEGMENTS = LEAN_DEFAULT, None, None


# (d_model, side, B'): 4-slot ring (few sequences) and 3-slot ring (more sequences than 7 CTAs per SM hold: 10 blocks x 27
# frames x 4 = 1080 > 1036); side 9 / 12 leave a partial last tile and a tail that straddles tiles in both directions
@pytest.mark.parametrize("chain", [0, 3])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(64, 12, 2), (320, 9, 27), (640, 36, 2), (320, 16, 3)])
def test_lean_kernel_is_bit_identical_to_the_general_kernel(shape, dtype, chain):
    d_model, side, Bp = shape
    kw, sd, x, id_emb, conds, masks, want = _case(d_model, side, Bp, dtype, "ones", 4242 + d_model + side)
    ours = _ours(kw, sd, dtype)
    lean = _run(ours, x, id_emb, conds, masks, True, chain)
    general = _run(ours, x, id_emb, conds, masks, False, chain)
    assert not torch.isnan(lean).any()
    assert torch.equal(lean, general), f"lean != general: max diff {(lean.float() - general.float()).abs().max().item():.3e}"
    close(lean, want, dtype, tol=LAYER_TOL, what=f"lean d_model {d_model} side {side} {dtype} chain {chain}")


@pytest.mark.parametrize("chain", [0, 3])
@pytest.mark.parametrize("shape", [(64, 12, 2), (320, 9, 27)])
def test_lean_kernel_power_path_is_bit_identical(shape, chain):
    """A_logs kept in fp32 (the S4D-real structure survives, mamba_layer.py:1476-1490): the POWER decay path of both kernels."""
    from actalker_b200 import _lib
    d_model, side, Bp = shape
    kw, sd, x, id_emb, conds, masks, _ = _case(d_model, side, Bp, torch.bfloat16, "ones", 4242 + d_model + side)
    ours = _ours(kw, sd, torch.bfloat16)
    with torch.no_grad():
        ours.audio_unit.A_logs.copy_(sd["audio_unit.A_logs"].cuda())
    assert ours.audio_unit.derived()["a_kind"] == _lib.ACTK_A_POWER
    lean = _run(ours, x, id_emb, conds, masks, True, chain)
    general = _run(ours, x, id_emb, conds, masks, False, chain)
    assert not torch.isnan(lean).any() and torch.equal(lean, general)


def test_lean_kernel_really_runs_and_mixed_masks_fall_back():
    """The scan launch of an all-ones call carries the fp32 B|C planes (so the C side can take the lean kernel); a call
    with one partial mask keeps the general kernel and the same results as with the lean route switched off."""
    from actalker_b200 import mamba_layer as ml
    kw, sd, x, id_emb, conds, masks, want = _case(320, 16, 3, torch.bfloat16, "ones", 4242 + 320 + 16)
    ours = _ours(kw, sd, torch.bfloat16)
    seen = {}
    lib = ml._lib.load()
    real = lib.actk_masked_scan_fwd

    class Spy:
        def __call__(self, args, stream):
            a = args._obj
            seen["bc32"] = [bool(a.br[i].bc32) for i in range(a.n_branches)]
            return real(args, stream)
    try:
        lib.actk_masked_scan_fwd = Spy()
        _run(ours, x, id_emb, conds, masks, True, None)
        assert seen["bc32"] == [True, True]
        rect = torch.zeros_like(masks[1])
        rect[:, :, 32:96, 16:80] = 1
        mixed = [masks[0], rect]
        a = _run(ours, x, id_emb, conds, mixed, True, None)
        assert seen["bc32"] == [True, False]
    finally:
        lib.actk_masked_scan_fwd = real
    b = _run(ours, x, id_emb, conds, mixed, False, None)
    assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cols", [32, 64])
def test_gemm_fp32_side_output_is_the_widened_rounded_result(dtype, cols):
    """actk_gemm_problem.c_f32: the first 32 / 64 output columns also leave the kernel as fp32 == float(c[:, :cols])
    (x_proj's B|C columns, mamba_layer.py:1521), for every row tile including a partial last one, grouped launches."""
    from actalker_b200 import gemm
    torch.manual_seed(7)
    probs, outs = [], []
    for M, K, N in [(1000, 640, 128), (70, 640, 128), (129, 96, 160)]:
        a = torch.randn(M, K, device="cuda").to(dtype)
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dtype)
        c = torch.full((M, N), float("nan"), device="cuda", dtype=dtype)
        f = torch.full((M, cols), float("nan"), device="cuda", dtype=torch.float32)
        probs.append(gemm.Problem(a, w, c, f32=f))
        outs.append((a, w, c, f))
    gemm.run(probs)
    for a, w, c, f in outs:
        ref = (a.float() @ w.float().t()).to(dtype)
        assert torch.allclose(c.float(), ref.float(), rtol=2e-2, atol=2e-2)
        assert torch.equal(f, c[:, :cols].float())
