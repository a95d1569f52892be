"""CPU restatement of the reference's live masked Mamba layer — TEST INFRASTRUCTURE ONLY.

What it restates (all in /root/reference/src/models/base/mamba_layer.py)
    HSCANS_dynamic            :142-184   scan-order helper ('sweep' == identity, 'scan' == boustrophedon)
    SS2D_Unit.__init__        :1394-1447 parameters x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds
    SS2D_Unit.dt_init/A_log_init/D_init :1450-1502  (the synthetic parameter distribution)
    SS2D_Unit.forward_core    :1505-1548 K=2 bidirectional scan
    SS2D_cond_v10.__init__    :1902-1953
    SS2D_cond_v10.forward     :1955-1986 mask -> index -> gather -> scan -> scatter -> add -> LN -> out_proj

The op graph is kept op-for-op (including the identity encode/decode scatter
passes and the flipped copy) so that it can be diffed against the reference;
only the two third-party calls are replaced by the restatements in this
package (`selective_scan_ref`, `downsample`).

PINNED: tests/golden/make_golden.py runs the real reference classes in the
build container and tests/test_oracle.py checks this file against those
committed outputs (same state dict, same inputs -> same outputs).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .mask_downsample import downsample
from .selective_scan_ref import selective_scan_ref

__all__ = ["HSCANS_dynamic_ref", "SS2D_Unit_ref", "SS2D_cond_v10_ref"]


class HSCANS_dynamic_ref:
    """mamba_layer.py:142-184.  Holds a permutation `order` of range(size) and its inverse;
    encode writes element j of the input to slot inv[j] (scatter), decode to slot order[j]."""

    def __init__(self, size: int, scan_type: str = "scan"):
        size = int(size)
        order = np.arange(size)
        if scan_type == "sweep":
            pass
        elif scan_type == "scan":
            grid = order.reshape(size, size)       # upstream raises here too unless size is a square count
            grid[1::2] = grid[1::2, ::-1]
            order = grid.reshape(-1)
        else:
            raise Exception("invalid encoder mode")
        self.index_flat = torch.from_numpy(order.astype(np.int64))[None, None]
        self.index_flat_inv = torch.from_numpy(np.argsort(order).astype(np.int64))[None, None]

    def to(self, device):
        self.index_flat = self.index_flat.to(device)
        self.index_flat_inv = self.index_flat_inv.to(device)
        return self

    def encode(self, img):
        return torch.zeros_like(img).scatter_(2, self.index_flat_inv.expand(img.shape), img)

    def decode(self, img):
        return torch.zeros_like(img).scatter_(2, self.index_flat.expand(img.shape), img)


class SS2D_Unit_ref(nn.Module):
    """mamba_layer.py:1394-1553.  Only the K=2 topology the live layer uses is restated."""

    def __init__(self, d_model, d_cond, cond_size=0, d_state=16, d_conv=3, expand=2, dt_rank="auto",
                 dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4,
                 dropout=0.0, conv_bias=True, bias=False, device=None, dtype=None, size=8,
                 scan_type="scan", num_direction=8, **kwargs):
        super().__init__()
        fk = {"device": device, "dtype": dtype}
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.d_cond = d_cond
        self.num_direction = K = num_direction
        self.scan_type = scan_type
        c = self.dt_rank + 2 * d_state
        self.x_proj_weight = nn.Parameter(torch.stack(
            [nn.Linear(self.d_inner, c, bias=False, **fk).weight for _ in range(K)], dim=0))
        ws, bs = [], []
        for _ in range(K):
            w, b = self.dt_init(self.dt_rank, self.d_inner, dt_scale, dt_init, dt_min, dt_max, dt_init_floor, **fk)
            ws.append(w)
            bs.append(b)
        self.dt_projs_weight = nn.Parameter(torch.stack(ws, dim=0))      # (K, D, R)
        self.dt_projs_bias = nn.Parameter(torch.stack(bs, dim=0))        # (K, D)
        a_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device))
        self.A_logs = nn.Parameter(a_log.repeat(K * self.d_inner, 1))     # (K*D, N)
        self.Ds = nn.Parameter(torch.ones(K * self.d_inner, dtype=torch.float32, device=device))
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else None

    @staticmethod
    def dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1,
                dt_init_floor=1e-4, **fk):
        lin = nn.Linear(dt_rank, d_inner, bias=True, **fk)
        std = dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(lin.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(lin.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(d_inner, **fk) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))        # softplus^-1
        with torch.no_grad():
            lin.bias.copy_(inv_dt)
        return lin.weight, lin.bias

    def forward_core(self, x: torch.Tensor, selective_scan=selective_scan_ref):
        Bsz, C, L = x.shape
        K = self.num_direction
        if K != 2:
            raise NotImplementedError("only num_direction=2 is live (TransformerSTmodel.py:3962-3971)")
        scans = HSCANS_dynamic_ref(size=L, scan_type=self.scan_type).to(x.device)
        xs = torch.stack([scans.encode(x.view(Bsz, -1, L))], dim=1).view(Bsz, K // 2, -1, L)
        xs = torch.cat([xs, torch.flip(xs, dims=[-1])], dim=1)                      # (B, K, D, L)
        x_dbl = torch.einsum("b k d l, k c d -> b k c l", xs.view(Bsz, K, -1, L), self.x_proj_weight)
        dts, Bs, Cs = torch.split(x_dbl, [self.dt_rank, self.d_state, self.d_state], dim=2)
        dts = torch.einsum("b k r l, k d r -> b k d l", dts.view(Bsz, K, -1, L), self.dt_projs_weight)
        xs = xs.view(Bsz, -1, L)
        dts = dts.contiguous().view(Bsz, -1, L)
        Bs = Bs.view(Bsz, K, -1, L)
        Cs = Cs.view(Bsz, K, -1, L)
        Ds = self.Ds.view(-1)
        As = -torch.exp(self.A_logs).view(-1, self.d_state)
        dt_bias = self.dt_projs_bias.view(-1)
        out_y = selective_scan(xs, dts, As, Bs, Cs, Ds, z=None, delta_bias=dt_bias,
                               delta_softplus=True, return_last_state=False).view(Bsz, K, -1, L)
        inv_y = torch.flip(out_y[:, K // 2:K], dims=[-1]).view(Bsz, K // 2, -1, L)
        return scans.decode(out_y[:, 0]) + scans.decode(inv_y[:, 0])

    def forward(self, x, selective_scan=selective_scan_ref):
        return self.forward_core(x, selective_scan)


class SS2D_cond_v10_ref(nn.Module):
    """mamba_layer.py:1902-1986."""

    def __init__(self, d_model, d_cond, cond_size=0, d_state=16, d_conv=3, expand=2, dt_rank="auto",
                 dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4,
                 dropout=0.0, conv_bias=True, bias=False, device=None, dtype=None, size=8,
                 scan_type="scan", num_direction=8, **kwargs):
        super().__init__()
        fk = {"device": device, "dtype": dtype}
        unit_args = (d_model, d_cond, cond_size, d_state, d_conv, expand, dt_rank, dt_min, dt_max, dt_init,
                     dt_scale, dt_init_floor, dropout, conv_bias, bias, device, dtype, size, scan_type,
                     num_direction)
        self.audio_unit = SS2D_Unit_ref(*unit_args)
        self.exp_unit = SS2D_Unit_ref(*unit_args)
        self.d_model, self.d_state, self.d_cond = d_model, d_state, d_cond
        self.d_inner = int(expand * d_model)
        self.audio_proj = nn.Linear(d_cond, self.d_inner, bias=bias, **fk)
        self.exp_proj = nn.Linear(d_cond, self.d_inner, bias=bias, **fk)
        self.id_proj = nn.Linear(d_cond, self.d_inner, bias=bias, **fk)
        self.in_proj1 = nn.Linear(d_model, self.d_inner, bias=bias, **fk)
        self.in_proj2 = nn.Linear(d_model, self.d_inner, bias=bias, **fk)
        self.out_norm = nn.LayerNorm(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)
        self.num_direction, self.scan_type = num_direction, scan_type

    def _branch(self, xz, mask, tail, unit, selective_scan):
        m = downsample(mask[:, 0, :, :], mask.shape[0], xz.shape[1], 1)
        idx = m.view(-1).int().nonzero().view(-1)
        sel = xz[:, idx, :]
        n = sel.shape[1]
        seq = torch.cat([sel] + tail, dim=1)
        out = unit(seq.permute(0, 2, 1), selective_scan).to(xz.dtype)
        xz[:, idx, :] = out[:, :, :n].permute(0, 2, 1)
        return xz

    def forward(self, x, id_emb, conds, masks, selective_scan=selective_scan_ref):
        audio_cond, exp_cond = conds[:, :-1], conds[:, -1:]
        id_tok = F.silu(self.id_proj(id_emb))
        xz1 = self._branch(self.in_proj1(x), masks[0], [id_tok, F.silu(self.audio_proj(audio_cond))],
                           self.audio_unit, selective_scan)
        xz2 = self._branch(self.in_proj2(x), masks[1], [id_tok, F.silu(self.exp_proj(exp_cond))],
                           self.exp_unit, selective_scan)
        return self.out_proj(self.out_norm(xz2 + xz1))
