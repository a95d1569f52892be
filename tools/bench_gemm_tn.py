"""The layer's projections on this repo's tcgen05 kernel (actk_gemm_tn_fwd) beside cuBLAS (torch) on the same operands.

For every product of one layer call at BASELINE configs[1] (B' = 25 x 72x72 tokens, d_model 320, bf16) and at the UNet's
other widths: time per launch (CUDA events, L2 flushed between launches), the bytes the product must move
(A + W + C once) and that figure against the measured HBM peak.  One JSON line per product.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from actalker_b200 import gemm  # noqa: E402

dev = "cuda"
dt = torch.bfloat16
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
peak = 6557.1
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])


def timed(fn, n=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1e3


def case(name, M, N, K, planes=1, lda=None, groups=1, silu=False):
    lda = lda or K
    As = [torch.randn(M, lda, device=dev, dtype=dt) for _ in range(groups)]
    Ws = [torch.randn(N, K, device=dev, dtype=dt) / K ** 0.5 for _ in range(groups)]
    a_views = [a[:, lda - K:] for a in As]
    outs = [torch.empty((planes, M, N // planes) if planes > 1 else (M, N), device=dev, dtype=dt) for _ in range(groups)]
    probs = [gemm.Problem(a, w, o, planes=planes) for a, w, o in zip(a_views, Ws, outs)]
    ours = timed(lambda: gemm.run(probs, silu=silu))

    def cublas():
        for a, w in zip(a_views, Ws):
            y = torch.nn.functional.linear(a, w)
            if silu:
                torch.nn.functional.silu(y)
    ref = timed(cublas)
    want = torch.nn.functional.linear(a_views[0], Ws[0])
    got = outs[0] if planes == 1 else torch.cat([outs[0][i] for i in range(planes)], dim=1)
    if silu:
        want = torch.nn.functional.silu(want)
    err = (got.float() - want.float()).abs().max().item()
    byts = groups * 2 * (M * K + N * K + M * N)
    print(json.dumps({"product": name, "M": M, "N": N, "K": K, "launch_problems": groups, "ours_us": round(ours, 1),
                      "cublas_us": round(ref, 1), "bytes": byts, "ours_GBps": round(byts / ours / 1e3, 1),
                      "frac_of_measured_hbm_peak": round(byts / ours / 1e3 / peak, 3), "max_abs_diff_vs_cublas": err}),
          flush=True)


if __name__ == "__main__":
    for d_model, side, Bp in [(320, 72, 25), (640, 36, 50), (1280, 18, 50)]:
        M, D = Bp * side * side, 2 * d_model
        R = d_model // 16
        rp = next(16 * k for k in (2, 3, 5) if R <= 16 * k)
        xw = 64 + 2 * rp
        tag = f"d_model {d_model}: "
        case(tag + "in_proj1|2 (stacked, 2 planes)", M, 2 * D, d_model, planes=2)
        case(tag + "x_proj, both branches", M, xw, D, groups=2)
        case(tag + "dt_proj (block-diagonal, dt columns in place), both branches", M, 2 * D, 2 * rp, lda=xw, groups=2)
        case(tag + "out_proj", M, d_model, D)
        case(tag + "audio_proj + SiLU (32 tokens per frame)", Bp * 33, D, 1024, silu=True)
