"""Small-shape pass over every kernel family in a few seconds, written as the target for compute-sanitizer:

    compute-sanitizer --tool memcheck|synccheck python tools/sanitize.py

(the shared B200 pool of this build refuses compute-sanitizer, so round 1 only ran it plain: every case below finishes
and returns finite values; out-of-range writes are looked for by the NaN-poisoned-output parity tests instead).

Covers the operator scan (dstate 16 and generic, ragged dims, z / last_state), the fused masked scan in its three launch
shapes (plain, chained chunks, two-level) over all-ones / rectangle / zero masks with both decay paths, partial channel
blocks (d_model 40 -> D = 80), the merge + LayerNorm kernels, the opt-in tcgen05 fusions and the sharded slice path.
Prints one line per case; results are also checked for finiteness so a silent fault cannot pass.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from actalker_b200 import SS2D_cond_v10, mamba_layer as ml, selective_scan_fn


def make_layer(dm, side, dtype, structured):
    torch.manual_seed(dm + side)
    layer = SS2D_cond_v10(d_model=dm, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=side,
                          scan_type="sweep", num_direction=2).eval()
    keep = {n: p.data.clone() for n, p in layer.named_parameters()}
    layer = layer.to(dtype)
    for n, p in layer.named_parameters():
        if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
            p.data = keep[n] if structured else p.data.float()
    return layer.cuda()


def masks(kind, dtype):
    px = 64
    if kind == "ones":
        return [torch.ones(1, 1, px, px, device="cuda", dtype=dtype)] * 2
    if kind == "zeros":
        return [torch.zeros(1, 1, px, px, device="cuda", dtype=dtype)] * 2
    a = torch.zeros(1, 1, px, px, device="cuda", dtype=dtype)
    b = torch.zeros(1, 1, px, px, device="cuda", dtype=dtype)
    a[..., 40:60, 12:50] = 1
    b[..., 4:36, 8:56] = 1
    return [a, b]


def layer_case(dm, side, Bp, dtype, kind, structured, seg=None, chain=None, tag=""):
    layer = make_layer(dm, side, dtype, structured)
    L = side * side
    x = torch.randn(Bp, L, dm, device="cuda").to(dtype)
    idm = torch.randn(Bp, 1, 1024, device="cuda").to(dtype)
    cd = torch.randn(Bp, 33, 1024, device="cuda").to(dtype)
    ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = seg, chain
    try:
        with torch.no_grad():
            y = layer(x, idm, cd, masks(kind, dtype))
        torch.cuda.synchronize()
    finally:
        ml.SCAN_SEGMENTS = ml.SCAN_CHAIN = None
    assert torch.isfinite(y.float()).all()
    print(f"layer d_model={dm} {side}x{side} B'={Bp} {str(dtype)[6:]} {kind} structured={structured} seg={seg} "
          f"chain={chain} {tag}: ok", flush=True)


def operator_case(batch, dim, L, N, dtype, z, last):
    g = torch.Generator().manual_seed(L)
    u = torch.randn(batch, dim, L, generator=g).to(dtype).cuda()
    delta = torch.randn(batch, dim, L, generator=g).to(dtype).cuda()
    A = -torch.exp(torch.randn(dim, N, generator=g)).cuda()
    B = torch.randn(batch, 2, N, L, generator=g).to(dtype).cuda()
    C = torch.randn(batch, 2, N, L, generator=g).to(dtype).cuda()
    D = torch.randn(dim, generator=g).cuda()
    bias = torch.randn(dim, generator=g).cuda()
    zz = torch.randn(batch, dim, L, generator=g).to(dtype).cuda() if z else None
    out = selective_scan_fn(u, delta, A, B, C, D, zz, bias, delta_softplus=True, return_last_state=last)
    torch.cuda.synchronize()
    y = out[0] if last else out
    assert torch.isfinite(y.float()).all()
    print(f"operator {batch}x{dim}x{L} N={N} {str(dtype)[6:]} z={z} last_state={last}: ok", flush=True)


def main():
    bf, f16, f32 = torch.bfloat16, torch.float16, torch.float32
    for dtype in (bf, f32):
        operator_case(2, 128, 197, 16, dtype, False, False)
        operator_case(1, 72, 33, 16, dtype, True, True)
    operator_case(2, 64, 50, 8, f16, False, False)
    operator_case(1, 64, 40, 24, f32, True, True)
    for kind in ("ones", "rect", "zeros"):
        layer_case(64, 8, 3, bf, kind, False)
    layer_case(64, 8, 3, f32, "rect", True)
    layer_case(40, 8, 2, f16, "rect", False, tag="(partial channel block)")
    layer_case(64, 16, 2, bf, "ones", False, chain=4, tag="(chained chunks)")
    layer_case(64, 16, 2, bf, "rect", True, chain=3, tag="(chained chunks)")
    layer_case(64, 16, 1, bf, "ones", False, seg=3, tag="(two-level)")
    layer_case(64, 16, 1, f32, "rect", False, seg=2, tag="(two-level)")
    # opt-in tensor-core fusions
    ml.FUSE_DT_PROJ = True
    try:
        layer_case(64, 8, 2, bf, "rect", False, tag="(fused dt_proj)")
        layer_case(64, 16, 2, f16, "ones", False, chain=3, tag="(fused dt_proj, chained)")
    finally:
        ml.FUSE_DT_PROJ = False
    ml.FUSE_LN_OUT_PROJ = True
    try:
        layer_case(320, 8, 3, bf, "rect", False, tag="(fused merge + LN + out_proj)")
    finally:
        ml.FUSE_LN_OUT_PROJ = False
    print("sanitize pass complete", flush=True)


if __name__ == "__main__":
    main()
