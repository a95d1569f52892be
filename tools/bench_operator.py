#!/usr/bin/env python
"""Time the operator-contract kernel (selective_scan_fn, channels-first) at the shapes the reference issues for
BASELINE config 2 (mamba_layer.py:1532-1538): u, delta (25, 1280, L'), B, C (25, 2, 16, L'), bf16."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from actalker_b200 import selective_scan_fn, _lib

def main():
    dev = "cuda"
    for dtype in (torch.bfloat16, torch.float32):
        for Lp in (5217, 5186):
            g = torch.Generator(device=dev).manual_seed(0)
            B, Dm = int(os.environ.get("OPB", "25")), 1280
            u = torch.randn(B, Dm, Lp, device=dev, generator=g).to(dtype)
            delta = torch.randn(B, Dm, Lp, device=dev, generator=g).to(dtype)
            A = -torch.exp(torch.randn(Dm, 16, device=dev, generator=g))
            Bm = torch.randn(B, 2, 16, Lp, device=dev, generator=g).to(dtype)
            Cm = torch.randn(B, 2, 16, Lp, device=dev, generator=g).to(dtype)
            Dv = torch.ones(Dm, device=dev); bias = torch.randn(Dm, device=dev) - 2
            for _ in range(3):
                selective_scan_fn(u, delta, A, Bm, Cm, Dv, None, bias, True)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); s.record()
            n = 10
            for _ in range(n):
                selective_scan_fn(u, delta, A, Bm, Cm, Dv, None, bias, True)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / n
            q = _lib.load().actk_scan_algorithmic_bytes(B, Lp, Dm, 2, 16, u.element_size())
            print(f"selective_scan_fn {str(dtype)[6:]:8s} L'={Lp}: {ms:.3f} ms  {q / ms / 1e6:.0f} GB/s algorithmic")

def vllm_kernel():
    """The upstream-derived CUDA kernel that ships in vLLM (mamba-ssm's selective_scan_fwd, built for this GPU): the
    recompiled reference kernel this repo's scan is measured against — same tensors, same contract."""
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn as vllm_scan
    except Exception as e:   # noqa: BLE001
        print("vllm kernel not importable:", type(e).__name__, str(e)[:200])
        return
    dev = "cuda"
    lib = _lib.load()
    for dtype in (torch.bfloat16, torch.float32):
        Lp, B, Dm = 5217, 25, 1280
        g = torch.Generator(device=dev).manual_seed(0)
        u = torch.randn(B, Dm, Lp, device=dev, generator=g).to(dtype)
        delta = torch.randn(B, Dm, Lp, device=dev, generator=g).to(dtype)
        A = -torch.exp(torch.randn(Dm, 16, device=dev, generator=g))
        Bm = torch.randn(B, 2, 16, Lp, device=dev, generator=g).to(dtype)
        Cm = torch.randn(B, 2, 16, Lp, device=dev, generator=g).to(dtype)
        Dv = torch.ones(Dm, device=dev); bias = torch.randn(Dm, device=dev) - 2
        state = torch.zeros(B, Dm, 16, device=dev, dtype=dtype)
        flag = torch.zeros(B, dtype=torch.bool, device=dev)
        work = delta.clone()                      # the kernel writes its output over delta
        ts = []
        try:
            for i in range(13):
                work.copy_(delta)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); vllm_scan(u, state, work, A, Bm, Cm, Dv, None, bias, True, has_initial_state=flag); e.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(s.elapsed_time(e))
        except Exception as e:   # noqa: BLE001
            print("vllm kernel not runnable:", type(e).__name__, str(e)[:200])
            return
        ms = sorted(ts)[len(ts) // 2]
        q = lib.actk_scan_algorithmic_bytes(B, Lp, Dm, 2, 16, u.element_size())
        print(f"vllm (mamba-ssm derived) selective_scan_fwd {str(dtype)[6:]:8s} L'={Lp}: {ms:.3f} ms  {q / ms / 1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
    vllm_kernel()
