"""GPU parity at the UNet's OTHER widths and at BASELINE configs[0]'s exact shape, against the CPU oracle.

The reference builds one masked Mamba layer per transformer block with `d_model = in_channels`
(src/models/base/TransformerSTmodel.py:3962-3971): 320 at 72x72 tokens, 640 at 36x36, 1280 at 18x18, so besides the
D = 640 layer the bench measures, every UNet call runs ten layers with D = 1280 / 2560, dt_rank 40 / 80 (x_dbl rows of
64 + 2*48 and 64 + 2*80 columns).  These tests drive those widths through the C-ABI in all three launch shapes of the
scan (single level, chained chunks, two-level) and compare with `oracle.SS2D_cond_v10_ref` (the restatement of
mamba_layer.py:1955-1986 pinned by the goldens) on the same seeded inputs.

Tolerances are the layer-level ones of test_gpu_parity.py (fp32 1e-3/1e-4 against the fp32 oracle, bf16 3e-2, fp16 5e-3).
"""
import pytest
import torch

from oracle import SS2D_cond_v10_ref
from test_gpu_parity import LAYER_TOL, close

pytestmark = pytest.mark.gpu

_FP32_PARAMS = ("A_logs", "Ds", "dt_projs_bias")   # Inference.py:430-433
_ORACLE = {}                                       # (d_model, dtype, mask kind) -> inputs, state dict, oracle output


def _keep_fp32(mod):
    for name, p in mod.named_parameters():
        if any(s in name for s in _FP32_PARAMS):
            p.data = p.data.float()


def _masks(kind, side, dtype):
    px = side * 8
    ones = torch.ones(1, 1, px, px)
    if kind == "ones":
        return [ones.to(dtype), ones.clone().to(dtype)]
    # the mouth / upper-face rectangle pair of SURVEY.md §8(d): disjoint regions of different sizes, token aligned
    mouth = torch.zeros(1, 1, px, px)
    mouth[:, :, (5 * side // 9) * 8:(8 * side // 9) * 8, (side // 3) * 8:(2 * side // 3) * 8] = 1
    upper = torch.zeros(1, 1, px, px)
    upper[:, :, (side // 9) * 8:(5 * side // 9) * 8, (side // 6) * 8:(5 * side // 6) * 8] = 1
    return [mouth.to(dtype), upper.to(dtype)]


def _case(d_model, side, Bp, dtype, mkind, seed):
    key = (d_model, side, Bp, dtype, mkind)
    if key in _ORACLE:
        return _ORACLE[key]
    torch.manual_seed(seed)
    kw = dict(d_model=d_model, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=side, scan_type="sweep",
              num_direction=2)
    ref = SS2D_cond_v10_ref(**kw).eval()
    with torch.no_grad():   # trained-like: general A in one branch, S4D structure kept in the other
        ref.exp_unit.A_logs.add_(0.5 * torch.randn_like(ref.exp_unit.A_logs))
        ref.audio_unit.Ds.copy_(1.0 + 0.2 * torch.randn_like(ref.audio_unit.Ds))
        ref.out_norm.weight.add_(0.1 * torch.randn_like(ref.out_norm.weight))
        ref.out_norm.bias.add_(0.1 * torch.randn_like(ref.out_norm.bias))
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    if dtype != torch.float32:
        ref = ref.to(dtype)
        _keep_fp32(ref)
    L = side * side
    x = torch.randn(Bp, L, d_model).to(dtype)
    id_emb = torch.randn(Bp, 1, 1024).to(dtype)
    conds = torch.randn(Bp, 33, 1024).to(dtype)
    masks = _masks(mkind, side, dtype)
    with torch.no_grad():
        want = ref(x.clone(), id_emb, conds, masks)
    _ORACLE[key] = (kw, sd, x, id_emb, conds, masks, want)
    return _ORACLE[key]


def _ours(kw, sd, dtype):
    from actalker_b200 import SS2D_cond_v10
    ours = SS2D_cond_v10(**kw).eval()
    ours.load_state_dict(sd, strict=True)
    if dtype != torch.float32:
        ours = ours.to(dtype)
        _keep_fp32(ours)
    return ours.cuda()


@pytest.mark.parametrize("launch", ["default", "chain", "two_level"])
@pytest.mark.parametrize("mkind", ["ones", "rects"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("width", [(640, 36, 2), (1280, 18, 4), (640, 3, 1)])   # last: 9 rows, odd: the split merge's row tail
def test_layer_matches_oracle_at_unet_widths(width, dtype, mkind, launch):
    """d_model 640 (36x36 tokens, D = 1280, dt_rank 40 -> rank pad 48) and 1280 (18x18, D = 2560, dt_rank 80) against the
    oracle, fp32 / bf16 / fp16, all-ones masks and a mouth / upper-face rectangle pair, three launch shapes."""
    from actalker_b200 import mamba_layer as ml
    d_model, side, Bp = width
    kw, sd, x, id_emb, conds, masks, want = _case(d_model, side, Bp, dtype, mkind, 72589 + d_model)
    ours = _ours(kw, sd, dtype)
    assert ours.audio_unit.dt_rank == d_model // 16 and ours.d_inner == 2 * d_model
    try:
        ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = {"default": (None, None), "chain": (1, 3), "two_level": (3, 0)}[launch]
        with torch.no_grad():
            got = ours(x.cuda(), id_emb.cuda(), conds.cuda(), [m.cuda() for m in masks])
    finally:
        ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = None, None
    assert got.shape == want.shape and got.dtype == dtype
    close(got, want, dtype, tol=LAYER_TOL, what=f"d_model {d_model} {dtype} {mkind} {launch}")


def test_baseline_config0_exact_shape_fp32():
    """BASELINE.json configs[0]: B=1, 14 frames x 32x32 latent tokens, d_model 320, d_state 16, 2 branches, fp32 —
    the case the reference can run on a CPU through selective_scan_ref.  Reference-initialised parameters
    (mamba_layer.py:1450-1502), all-ones masks (Inference.py:545-546), L' = 1057 / 1026."""
    kw, sd, x, id_emb, conds, masks, want = _case(320, 32, 14, torch.float32, "ones", 72589)
    ours = _ours(kw, sd, torch.float32)
    with torch.no_grad():
        got = ours(x.cuda(), id_emb.cuda(), conds.cuda(), [m.cuda() for m in masks])
    m0 = ours.mask_cache.get(masks[0].cuda(), 1024)
    assert m0.n_sel == 1024 and got.shape == (14, 1024, 320)
    close(got, want, torch.float32, tol=LAYER_TOL, what="configs[0] 14 x 32x32 fp32")


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_config1_full_width_frames_match_oracle(dtype):
    """Frames of BASELINE configs[1]'s exact per-frame shape (72x72 tokens, d_model 320, bf16, all-ones masks, L' = 5217 /
    5186) against the CPU oracle — the full-size parity the bench line's `parity` field repeats on its own frames.
    Two frames keep the sequential Python token loop of the oracle within seconds."""
    kw, sd, x, id_emb, conds, masks, want = _case(320, 72, 2, dtype, "ones", 72589 + 1)
    ours = _ours(kw, sd, dtype)
    with torch.no_grad():
        got = ours(x.cuda(), id_emb.cuda(), conds.cuda(), [m.cuda() for m in masks])
    close(got, want, dtype, tol=LAYER_TOL, what="configs[1] frames")


def test_masks_on_the_cpu_work_like_upstream():
    """The reference indexes `xz[:, idx, :]` with whatever device the mask lives on (indexing moves the index tensor); here
    a host mask must give the same layer output as the same mask on the GPU — never a host pointer in a kernel."""
    kw, sd, x, id_emb, conds, masks, want = _case(640, 36, 2, torch.float32, "rects", 72589 + 640)
    ours = _ours(kw, sd, torch.float32)
    with torch.no_grad():
        on_gpu = ours(x.cuda(), id_emb.cuda(), conds.cuda(), [m.cuda() for m in masks])
        on_cpu = ours(x.cuda(), id_emb.cuda(), conds.cuda(), masks)
    assert torch.equal(on_gpu, on_cpu)
    close(on_cpu, want, torch.float32, tol=LAYER_TOL, what="cpu masks")
