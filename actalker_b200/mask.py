"""Region mask -> token index list, the integer half of the hot path.

Reference: src/models/base/mamba_layer.py:1962-1963 (and :1973-1974)

    m   = IPAdapterMaskProcessor.downsample(masks[i][:, 0, :, :], masks[i].shape[0], L, 1)
    idx = m.view(-1).int().nonzero().view(-1)

`downsample` is diffusers 0.29.2's (requirements.txt:10): a bicubic `F.interpolate` of the pixel mask to the
token grid.  It stays in PyTorch with the identical call so the values — and therefore the truncating
`.int()` — are bit-identical to what the reference computes on the same device and dtype.  What changes is
WHEN it runs: the reference recomputes it (with two host syncs) in every layer call; masks are constant for
a whole clip (pipeline ...two_ip.py:702-711), so the index list is cached per (mask storage, version, L) and
the hot path sees no `nonzero()` sync after the first call.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

__all__ = ["downsample", "mask_to_index", "MaskIndex", "MaskIndexCache"]


def downsample(mask: torch.Tensor, batch_size: int, num_queries: int, value_embed_dim: int) -> torch.Tensor:
    """(b, H, W) mask -> (batch_size, num_queries, value_embed_dim); same arithmetic as diffusers 0.29.2."""
    o_h, o_w = mask.shape[1], mask.shape[2]
    ratio = o_w / o_h
    mask_h = int(math.sqrt(num_queries / ratio))
    mask_h = int(mask_h) + int((num_queries % int(mask_h)) != 0)
    mask_w = num_queries // mask_h
    m = F.interpolate(mask.unsqueeze(0), size=(mask_h, mask_w), mode="bicubic").squeeze(0)
    if m.shape[0] < batch_size:
        m = m.repeat(batch_size, 1, 1)
    m = m.view(m.shape[0], -1)
    n_down = mask_h * mask_w
    if n_down < num_queries:
        m = F.pad(m, (0, num_queries - m.shape[1]), value=0.0)
    if n_down > num_queries:
        m = m[:, :num_queries]
    return m.view(m.shape[0], m.shape[1], 1).repeat(1, 1, value_embed_dim)


def mask_to_index(mask: torch.Tensor, num_queries: int) -> torch.Tensor:
    """(b, 1, H, W) -> ascending int64 token indices, exactly mamba_layer.py:1962-1963."""
    m = downsample(mask[:, 0, :, :], mask.shape[0], num_queries, 1)
    return m.view(-1).int().nonzero().view(-1)


@dataclass
class MaskIndex:
    idx: torch.Tensor        # (n_sel,) int32, ascending, on the LAYER's device — what the kernels read
    idx64: torch.Tensor      # (n_sel,) int64 — for torch-side gathers of the small x_dbl rows
    selected: torch.Tensor   # (L,) uint8 row flags for the merge kernel
    n_sel: int               # host copy (one sync when the entry is created)
    L: int
    mask_ref: torch.Tensor = None   # keeps the keyed storage alive so its address cannot be recycled
    weight: torch.Tensor = None     # (L,) downsampled mask values in the mask's dtype (v8/v9 multiplicative blend)

    @property
    def full(self) -> bool:
        return self.n_sel == self.L


class MaskIndexCache:
    """Per-layer cache keyed by the mask tensor's storage, version counter, shape, dtype and L."""

    def __init__(self, max_entries: int = 16):
        self._entries: Dict[Tuple, MaskIndex] = {}
        self._max = max_entries
        self.misses = 0

    def get(self, mask: torch.Tensor, L: int, device=None) -> MaskIndex:
        """`device`: where the kernels will read the index list (the activations' device).  The downsample itself runs
        where the reference runs it — on the mask's own device, in the mask's dtype — and only its integer results are
        moved, as the reference's `xz[:, idx, :]` does implicitly when the mask lives on the CPU or on another GPU.
        Handing a host pointer to the kernels instead would fault the context (no Python exception)."""
        device = mask.device if device is None else torch.device(device)
        key = (mask.data_ptr(), mask._version, tuple(mask.shape), mask.dtype, str(mask.device), L, str(device))
        hit = self._entries.get(key)
        if hit is not None:
            return hit
        if mask.dim() != 4:
            raise RuntimeError(f"mask must be (b, 1, H, W) as pipeline ...two_ip.py:632-633 builds it, got {tuple(mask.shape)}")
        self.misses += 1
        down = downsample(mask[:, 0, :, :], mask.shape[0], L, 1)     # (b, L, 1), the reference's own expression
        idx64 = down.view(-1).int().nonzero().view(-1)
        if idx64.numel() and int(idx64[-1]) >= L:
            # masks with batch > 1 flatten to b*L entries upstream and then index out of range (:1963);
            # the live pipeline always passes batch 1 (pipeline ...two_ip.py:632-633).
            raise RuntimeError(f"mask of shape {tuple(mask.shape)} selects token {int(idx64[-1])} >= L={L}")
        idx64 = idx64.to(device)
        sel = torch.zeros(L, dtype=torch.uint8, device=device)
        sel[idx64] = 1
        entry = MaskIndex(idx=idx64.to(torch.int32), idx64=idx64, selected=sel, n_sel=int(idx64.numel()), L=L,
                          mask_ref=mask, weight=down[0, :, 0].contiguous().to(device))
        if len(self._entries) >= self._max:
            self._entries.pop(next(iter(self._entries)))
        self._entries[key] = entry
        return entry
