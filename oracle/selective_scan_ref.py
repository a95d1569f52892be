"""CPU restatement of mamba-ssm's ``selective_scan_ref`` — TEST INFRASTRUCTURE ONLY.

What it restates
    ``mamba_ssm.ops.selective_scan_interface.selective_scan_ref`` of
    mamba-ssm==1.2.0.post1 (third-party; pinned by the reference at
    environment.yaml:43 and install_actalker.sh:93; imported at
    src/models/base/mamba_layer.py:21-23; live call site
    src/models/base/mamba_layer.py:1532-1538).  The package is NOT under
    /root/reference and cannot be installed in the build container, so this is
    a restatement of its published semantics:

        delta = softplus(delta + delta_bias)            (flags permitting)
        h_l   = exp(delta_l * A) * h_{l-1} + delta_l * B_l * u_l
        y_l   = <h_l, C_l> + D * u_l ;  out = y * silu(z)  (z optional)

    all arithmetic in float32 (or ``compute_dtype``), result cast to u.dtype.

PARITY UNPINNED by upstream tests (the reference has none for this path).
Cross-checks that do exist: tests/test_oracle.py compares this function with
transformers' independent ``MambaMixer.slow_forward`` recurrence, and
tests/golden/ holds outputs of the real reference layer driven through it.

The time loop is the same sequential recurrence as the upstream function; the
discretised tensors are produced one L-block at a time so the (B, D, L, N)
intermediates upstream materialises (~1.2 GB each at BASELINE config 1) stay
bounded.  Elementwise results are the same fp32 operations.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["selective_scan_ref"]


def _expand_groups(M: torch.Tensor, dim: int) -> torch.Tensor:
    """(B, G, N, L) -> (B, dim, N, L): channel d reads group d // (dim // G)."""
    G = M.shape[1]
    if dim % G != 0:
        raise ValueError(f"dim={dim} is not a multiple of groups={G}")
    return M.repeat_interleave(dim // G, dim=1)


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                       delta_softplus=False, return_last_state=False,
                       compute_dtype: torch.dtype = torch.float32,
                       l_block: int = 128):
    """u, delta: (B, Dm, L); A: (Dm, N) real; B, C: (Dm, N) | (B, N, L) | (B, G, N, L);
    D, delta_bias: (Dm,) or None; z: (B, Dm, L) or None.  Returns (B, Dm, L) in u.dtype
    (and the last state (B, Dm, N) when ``return_last_state``)."""
    if A.is_complex():
        raise NotImplementedError("complex A is outside the ACTalker hot path")
    dtype_in = u.dtype
    cd = compute_dtype
    u_f = u.to(cd)
    dl = delta.to(cd)
    if delta_bias is not None:
        dl = dl + delta_bias.to(cd)[..., None]
    if delta_softplus:
        dl = F.softplus(dl)  # torch: x if x > 20 else log1p(exp(x))
    batch, dim, L = u_f.shape
    N = A.shape[1]
    A_f = A.to(cd)
    var_B, var_C = B.dim() >= 3, C.dim() >= 3
    B_f, C_f = B.to(cd), C.to(cd)
    if var_B:
        B_f = B_f[:, None] if B_f.dim() == 3 else B_f      # (B, G, N, L)
        B_f = _expand_groups(B_f, dim) if B_f.shape[1] != 1 else B_f
    if var_C:
        C_f = C_f[:, None] if C_f.dim() == 3 else C_f
        C_f = _expand_groups(C_f, dim) if C_f.shape[1] != 1 else C_f

    h = torch.zeros(batch, dim, N, dtype=cd, device=u.device)
    y = torch.empty(batch, dim, L, dtype=cd, device=u.device)
    for l0 in range(0, L, l_block):
        l1 = min(L, l0 + l_block)
        d_blk = dl[:, :, l0:l1]                                  # (B, Dm, T)
        dA = torch.exp(d_blk[..., None] * A_f[None, :, None, :])  # (B, Dm, T, N)
        du = d_blk * u_f[:, :, l0:l1]                            # (B, Dm, T)
        if var_B:
            dBu = du[..., None] * B_f[:, :, :, l0:l1].transpose(2, 3)  # (B, Dm|1, T, N) bcast
        else:
            dBu = du[..., None] * B_f[None, :, None, :]
        for t in range(l1 - l0):
            h = dA[:, :, t] * h + dBu[:, :, t]
            if var_C:
                y[:, :, l0 + t] = (h * C_f[:, :, :, l0 + t]).sum(-1)
            else:
                y[:, :, l0 + t] = (h * C_f[None]).sum(-1)
    out = y if D is None else y + u_f * D.to(cd)[None, :, None]
    if z is not None:
        out = out * F.silu(z.to(cd))
    out = out.to(dtype_in)
    return (out, h) if return_last_state else out
