"""Multi-GPU execution of the masked selective-scan layer: one process per GPU, torch.distributed (NCCL).

The reference is single-GPU (no collective call sites anywhere).  Every (batch, branch,
direction, channel) recurrence of the layer is independent (mamba_layer.py:1532-1538), so the path shards two
ways:

  mode="batch"    the batch x CFG x frame axis B' is split across ranks.  No data-path collective at all; this
                  is how the denoising loop's frame windows / CFG copies spread over a box (weak scaling).
  mode="channel"  every rank runs the dense front half (in_proj, x_proj: they contract over ALL d_inner channels
                  and are cheap) and scans only its d_inner slice [lo, hi) of both branches and both directions
                  (B|C and the index list are replicated, delta / A / D / dt_bias / u are sliced).  LayerNorm
                  (mamba_layer.py:1984) needs complete channels, so the merged slices are exchanged with ONE
                  NCCL all-gather over NVLink, after which a LayerNorm kernel reads the gathered (rank, row,
                  slice) layout directly and out_proj follows (strong scaling of a single layer call).

`ShardPlan` is pure host logic and is what the world_size-2 gloo tests exercise on CPU; the compute itself has
no CPU path.
"""
from __future__ import annotations

import ctypes as ct
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .mamba_layer import SS2D_cond_v10, _timed
from .selective_scan_interface import _DTYPES, _ptr, _stream

__all__ = ["ShardPlan", "ShardedSS2DCondV10", "all_gather_slices"]


@dataclass(frozen=True)
class ShardPlan:
    """Partition of B' (batch mode) or of d_inner (channel mode) over `world` ranks."""
    mode: str
    world: int
    extent: int          # B' or d_inner
    granule: int = 1     # channel mode: slices are multiples of 8 channels (16-byte rows), equal on every rank

    def __post_init__(self):
        if self.mode not in ("batch", "channel"):
            raise ValueError("mode must be 'batch' or 'channel'")
        if self.world < 1 or self.extent < 1:
            raise ValueError("world and extent must be positive")
        if self.mode == "channel" and self.extent % (self.world * self.granule) != 0:
            raise ValueError(f"d_inner={self.extent} must be a multiple of world*{self.granule}={self.world * self.granule} "
                             "(equal 16-byte-granular channel slices; the all-gather needs equal contributions)")

    def bounds(self, rank: int) -> Tuple[int, int]:
        """[lo, hi) owned by `rank`.  Batch mode spreads the remainder over the first ranks."""
        if not 0 <= rank < self.world:
            raise ValueError("rank out of range")
        if self.mode == "channel":
            w = self.extent // self.world
            return rank * w, (rank + 1) * w
        base, rem = divmod(self.extent, self.world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)

    def all_bounds(self) -> List[Tuple[int, int]]:
        return [self.bounds(r) for r in range(self.world)]


def all_gather_slices(local: torch.Tensor, group=None) -> torch.Tensor:
    """(rows..., Ds) on every rank -> (world, rows..., Ds), rank-major — exactly what one all-gather writes, so no
    transposing copy is made; consumers index the leading rank axis (actk_gathered_layernorm_fwd does)."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


class ShardedSS2DCondV10(torch.nn.Module):
    """Wraps a (replicated) SS2D_cond_v10.  forward takes the same arguments as the layer.
    batch mode: each rank passes ITS rows of B' (the caller shards the batch) and gets its rows back.
    channel mode: every rank passes the same full inputs and gets the full output."""

    def __init__(self, layer: SS2D_cond_v10, mode: str = "channel", group=None):
        super().__init__()
        self.layer, self.mode, self.group = layer, mode, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.plan = ShardPlan("channel", self.world, layer.d_inner, 8) if mode == "channel" else None

    def forward(self, x, id_emb, conds, masks):
        layer = self.layer
        if self.mode == "batch" or self.world == 1:
            return layer(x, id_emb, conds, masks)
        lo, hi = self.plan.bounds(self.rank)
        xz1, xz2, tail1, tail2, m1, m2 = layer.project_inputs(x, id_emb, conds, masks)
        merged = layer.scan_core(xz1, xz2, tail1, tail2, m1, m2, ch_slice=(lo, hi))      # (B', L, Ds)
        with _timed("all_gather", x.device):
            gathered = all_gather_slices(merged, self.group)                               # (P, B', L, Ds)
        y = self.gathered_layernorm(gathered)
        return layer.out_proj(y)

    def gathered_layernorm(self, gathered: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        P, Bp, L, Ds = gathered.shape
        norm = self.layer.out_norm
        out = torch.empty((Bp, L, P * Ds), dtype=gathered.dtype, device=gathered.device)
        gamma, beta = norm.weight.to(gathered.dtype), norm.bias.to(gathered.dtype)
        with torch.cuda.device(gathered.device), _timed("gathered_ln", gathered.device):
            _lib.check(lib.actk_gathered_layernorm_fwd(_ptr(gathered), P, Bp * L, Ds, _ptr(gamma), _ptr(beta),
                                                       float(norm.eps), _ptr(out), _DTYPES[gathered.dtype],
                                                       _stream(gathered)), "actk_gathered_layernorm_fwd")
        return out
