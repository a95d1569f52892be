"""Host side of the tensor-core projections (C-ABI actk_gemm_tn_fwd, csrc/gemm_tn.cu).

Every dense contraction of the layer — in_proj1/2, the id / audio / expression token projections, x_proj, dt_proj and
out_proj (reference src/models/base/mamba_layer.py:1960-1961, :1966, :1972, :1977, :1521-1523, :1985) — is
`C = A @ W^T` with W stored as nn.Linear stores it.  With 16-bit activations they run on this repo's persistent
TMA + tcgen05 kernel, up to four independent products per launch; fp32 activations keep torch's fp32 GEMM (the
reference's own arithmetic for that dtype — the tensor cores would round the operands).
"""
from __future__ import annotations

import ctypes as ct
from typing import List, Optional, Sequence

import torch

from . import _lib

__all__ = ["Problem", "usable", "run", "linear"]

_DTYPES = {torch.float16: _lib.ACTK_F16, torch.bfloat16: _lib.ACTK_BF16}


class Problem:
    """C = A @ W^T.  a: (M, K) view with unit column stride; w: (N, K) with unit column stride;
    out: (M, N) view with unit column stride, or (planes, M, N / planes) contiguous planes."""

    __slots__ = ("a", "w", "out", "planes", "silu", "peers", "ldc", "f32")

    def __init__(self, a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, planes: int = 1, silu: bool = False,
                 peers: Sequence[int] = (), ldc: int = 0, f32: Optional[torch.Tensor] = None):
        """peers: device addresses of the (M, N) output slot in every rank's gather buffer (fused GEMM + all-gather over
        NVLink peer memory; `out` is then None and `ldc` the row pitch of those slots, N when 0).
        f32: optional (M, 32 | 64) fp32 tensor that also receives the first columns of the result, widened from the
        rounded 16-bit values (x_proj's B|C columns for the lean scan kernel)."""
        self.a, self.w, self.out, self.planes, self.silu, self.peers, self.ldc = a, w, out, planes, silu, tuple(peers), ldc
        self.f32 = f32

    def fill(self, p: "_lib.GemmProblem"):
        a, w, out = self.a, self.w, self.out
        M, K = a.shape
        N = w.shape[0]
        if a.dim() != 2 or w.dim() != 2 or w.shape[1] != K or a.stride(1) != 1 or w.stride(1) != 1:
            raise RuntimeError(f"gemm: a {tuple(a.shape)}/{a.stride()} and w {tuple(w.shape)}/{w.stride()} must be 2-D with unit column stride")
        if self.peers:
            if out is not None or self.planes != 1:
                raise RuntimeError("gemm: a problem with peer buffers has no local output tensor and one plane")
            ldc, plane_stride = self.ldc or N, 0
        elif self.planes == 1:
            if tuple(out.shape) != (M, N) or out.stride(1) != 1:
                raise RuntimeError(f"gemm: out {tuple(out.shape)}/{out.stride()} must be ({M}, {N}) with unit column stride")
            ldc, plane_stride = out.stride(0), 0
        else:
            if tuple(out.shape) != (self.planes, M, N // self.planes) or not out.is_contiguous():
                raise RuntimeError(f"gemm: out {tuple(out.shape)} must be contiguous ({self.planes}, {M}, {N // self.planes})")
            ldc, plane_stride = out.stride(1), out.stride(0)
        # a one-row operand may carry any row stride; the kernel only needs a pitch that satisfies its checks
        p.a, p.w, p.c = a.data_ptr(), w.data_ptr(), (None if self.peers else out.data_ptr())
        p.n_peers = len(self.peers)
        for i, ptr in enumerate(self.peers):
            p.peer_c[i] = ptr
        p.lda = a.stride(0) if M > 1 else K
        p.ldw = w.stride(0) if N > 1 else K
        p.ldc = ldc if M > 1 else N // self.planes
        p.plane_stride = plane_stride
        p.M, p.N, p.K, p.planes = M, N, K, self.planes
        p.epilogue = _lib.GEMM_EPI_SILU if self.silu else _lib.GEMM_EPI_NONE
        if self.f32 is not None:
            f = self.f32
            if f.dtype != torch.float32 or f.dim() != 2 or f.shape[0] != M or f.shape[1] not in (32, 64) or f.stride(1) != 1:
                raise RuntimeError(f"gemm: f32 side output {tuple(f.shape)} {f.dtype} must be ({M}, 32 | 64) float32")
            p.c_f32, p.ldc_f32, p.f32_cols = f.data_ptr(), (f.stride(0) if M > 1 else f.shape[1]), f.shape[1]
        else:
            p.c_f32, p.ldc_f32, p.f32_cols = None, 0, 0


def usable(*tensors: Optional[torch.Tensor]) -> bool:
    """True when every given tensor is a CUDA f16 / bf16 tensor of one dtype: the route this kernel serves."""
    ts = [t for t in tensors if t is not None]
    return bool(ts) and all(t.is_cuda and t.dtype in _DTYPES and t.dtype == ts[0].dtype for t in ts)


def run(problems: Sequence[Problem], silu: bool = False, name: str = "gemm"):
    """Launch the problems (groups of up to four per launch) on the current stream of their device.
    silu=True sets the SiLU epilogue on every problem (a Problem can also carry its own flag)."""
    from .mamba_layer import _timed   # timing hook shared with the other C-ABI launches
    lib = _lib.load()
    problems = [p for p in problems if p.a.shape[0] > 0]
    if not problems:
        return
    dev = problems[0].a.device
    dtype = _DTYPES[problems[0].a.dtype]
    if silu:
        for p in problems:
            p.silu = True
    stream = ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for i in range(0, len(problems), _lib.GEMM_MAX_PROBLEMS):
        group = problems[i:i + _lib.GEMM_MAX_PROBLEMS]
        arr = (_lib.GemmProblem * len(group))()
        for p, slot in zip(group, arr):
            p.fill(slot)
        with torch.cuda.device(dev), _timed(name, dev):
            _lib.check(lib.actk_gemm_tn_fwd(arr, len(group), dtype, stream), "actk_gemm_tn_fwd")


def linear(x: torch.Tensor, w: torch.Tensor, silu: bool = False, name: str = "gemm", out: torch.Tensor = None) -> torch.Tensor:
    """nn.Linear(bias=False) [+ SiLU] on (..., K) activations: this repo's kernel for 16-bit CUDA tensors, torch otherwise.
    out: optional contiguous (..., N) destination of x's dtype."""
    shape = (*x.shape[:-1], w.shape[0])
    if out is not None and (tuple(out.shape) != shape or not out.is_contiguous() or out.dtype != x.dtype):
        raise RuntimeError(f"gemm.linear: out {tuple(out.shape)} {out.dtype} must be contiguous {shape} {x.dtype}")
    if usable(x, w) and x.shape[-1] % 8 == 0 and w.shape[0] % 8 == 0 and (out is None or out.data_ptr() % 16 == 0):
        a = x.reshape(-1, x.shape[-1])
        if a.stride(1) != 1 or a.stride(0) % 8 or a.data_ptr() % 16:
            a = a.contiguous()
        wc = w if (w.stride(1) == 1 and w.stride(0) % 8 == 0 and w.data_ptr() % 16 == 0) else w.contiguous()
        if out is None:
            out = torch.empty(shape, dtype=x.dtype, device=x.device)
        run([Problem(a, wc, out.view(a.shape[0], w.shape[0]))], silu=silu, name=name)
        return out
    y = torch.nn.functional.linear(x, w.to(x.dtype))
    y = torch.nn.functional.silu(y) if silu else y
    if out is not None:
        out.copy_(y)
        return out
    return y
