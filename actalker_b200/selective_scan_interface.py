"""`selective_scan_fn` with mamba-ssm's signature, running on the sm_100a C-ABI library.

Mirrors `mamba_ssm.ops.selective_scan_interface.selective_scan_fn` (mamba-ssm 1.2.0.post1; imported by the
reference at src/models/base/mamba_layer.py:21-23, called at :1532-1538) for the argument forms the reference
uses: real A (dim, dstate), input-dependent B/C of shape (batch, dstate, L) or (batch, groups, dstate, L),
optional D / z / delta_bias, delta_softplus, return_last_state.  Forward only (the reference's inference path
runs under torch.no_grad(), pipeline ...two_ip.py:351).

`MAMBA_AVAILABLE` mirrors the reference's availability flag (mamba_layer.py:21-34).  There is no fallback:
without the compiled library or on a non-CUDA tensor the call raises.
"""
from __future__ import annotations

import ctypes as ct

import torch

from . import _lib

__all__ = ["selective_scan_fn", "MAMBA_AVAILABLE", "a_kind_of"]

_DTYPES = {torch.float32: _lib.ACTK_F32, torch.float16: _lib.ACTK_F16, torch.bfloat16: _lib.ACTK_BF16}

try:
    _lib.load()
    MAMBA_AVAILABLE = True
except _lib.LibraryMissing:
    MAMBA_AVAILABLE = False


def _ptr(t):
    return None if t is None else ct.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return ct.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def a_kind_of(A: torch.Tensor, rel_tol: float = 1e-6) -> int:
    """Probe A (dim, dstate) fp32 on its device: ACTK_A_POWER if A[d, n] == (n+1)*A[d, 0] (the S4D-real
    initialisation, mamba_layer.py:1476-1490), else ACTK_A_GENERAL.  One device->host read; callers cache it."""
    lib = _lib.load()
    if not A.is_cuda or A.dtype != torch.float32:
        raise RuntimeError("a_kind_of: A must be a CUDA float32 tensor")
    A = A.contiguous()
    flag = torch.empty(1, dtype=torch.int32, device=A.device)
    with torch.cuda.device(A.device):
        _lib.check(lib.actk_a_structure(_ptr(A), A.shape[0], A.shape[1], rel_tol, _ptr(flag), _stream(A)),
                   "actk_a_structure")
    return int(flag.item())


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False, a_kind="general"):
    """u, delta: (batch, dim, L); A: (dim, dstate); B, C: (batch, [groups,] dstate, L); D, delta_bias: (dim,);
    z: (batch, dim, L).  Returns (batch, dim, L) in u.dtype [and last_state (batch, dim, dstate) fp32].

    `a_kind` is an extension: "general" (default, always valid), "power" (caller guarantees the S4D-real
    structure) or "auto" (probe A on the device; costs one host sync)."""
    lib = _lib.load()
    if not u.is_cuda:
        raise RuntimeError("selective_scan_fn: tensors must live on a CUDA device (no CPU path in actalker_b200)")
    if u.dtype not in _DTYPES:
        raise RuntimeError(f"selective_scan_fn: unsupported dtype {u.dtype}")
    if A.is_complex():
        raise NotImplementedError("selective_scan_fn: complex A is not used by ACTalker and not built")
    if B.dim() < 3 or C.dim() < 3:
        raise NotImplementedError("selective_scan_fn: input-independent (dim, dstate) B/C is not used by ACTalker")
    if u.dim() != 3 or delta.shape != u.shape:
        raise RuntimeError(f"selective_scan_fn: u {tuple(u.shape)} / delta {tuple(delta.shape)} must be equal 3-D shapes")
    batch, dim, L = u.shape
    dstate = A.shape[1]
    if A.shape[0] != dim:
        raise RuntimeError(f"selective_scan_fn: A {tuple(A.shape)} does not match dim={dim}")
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    groups = B.shape[1]
    if B.shape != (batch, groups, dstate, L) or C.shape != B.shape:
        raise RuntimeError(f"selective_scan_fn: B {tuple(B.shape)} / C {tuple(C.shape)} must be (batch, groups, dstate, L)")
    if dim % groups != 0:
        raise RuntimeError(f"selective_scan_fn: dim={dim} not divisible by groups={groups}")

    def last_contig(t):
        return t if t.stride(-1) == 1 else t.contiguous()

    dt = u.dtype
    u, delta = last_contig(u), last_contig(delta.to(dt))
    B, C = last_contig(B.to(dt)), last_contig(C.to(dt))
    if z is not None:
        if z.shape != u.shape:
            raise RuntimeError("selective_scan_fn: z must have the shape of u")
        z = last_contig(z.to(dt))
    A = A.to(torch.float32).contiguous()
    D = None if D is None else D.to(torch.float32).contiguous()
    delta_bias = None if delta_bias is None else delta_bias.to(torch.float32).contiguous()
    for name, t in (("D", D), ("delta_bias", delta_bias)):
        if t is not None and t.shape != (dim,):
            raise RuntimeError(f"selective_scan_fn: {name} must have shape ({dim},)")
    if isinstance(a_kind, str):
        a_kind = {"general": _lib.ACTK_A_GENERAL, "power": _lib.ACTK_A_POWER}.get(a_kind) \
            if a_kind != "auto" else (a_kind_of(A) if dstate == 16 else _lib.ACTK_A_GENERAL)
        if a_kind is None:
            raise ValueError("a_kind must be 'general', 'power' or 'auto'")

    out = torch.empty((batch, dim, L), dtype=dt, device=u.device)
    last = torch.empty((batch, dim, dstate), dtype=torch.float32, device=u.device) if return_last_state else None
    if batch == 0 or L == 0 or dim == 0:
        return (out, last) if return_last_state else out
    a = _lib.ScanArgs()
    a.u, a.delta, a.B, a.C, a.z = _ptr(u), _ptr(delta), _ptr(B), _ptr(C), _ptr(z)
    a.A, a.D, a.delta_bias = _ptr(A), _ptr(D), _ptr(delta_bias)
    a.out, a.last_state = _ptr(out), _ptr(last)
    a.batch, a.dim, a.groups, a.dstate, a.seqlen = batch, dim, groups, dstate, L
    a.u_sb, a.u_sd = u.stride(0), u.stride(1)
    a.delta_sb, a.delta_sd = delta.stride(0), delta.stride(1)
    if z is not None:
        a.z_sb, a.z_sd = z.stride(0), z.stride(1)
    a.out_sb, a.out_sd = out.stride(0), out.stride(1)
    a.B_sb, a.B_sg, a.B_sn = B.stride(0), B.stride(1), B.stride(2)
    a.C_sb, a.C_sg, a.C_sn = C.stride(0), C.stride(1), C.stride(2)
    a.dtype, a.delta_softplus, a.a_kind = _DTYPES[dt], int(bool(delta_softplus)), int(a_kind)
    with torch.cuda.device(u.device):
        _lib.check(lib.actk_selective_scan_fwd(ct.byref(a), _stream(u)), "actk_selective_scan_fwd")
    return (out, last) if return_last_state else out
