"""CPU restatement of diffusers' ``IPAdapterMaskProcessor.downsample`` and of the
reference's mask -> token-index rule — TEST INFRASTRUCTURE ONLY.

What it restates
    ``diffusers.image_processor.IPAdapterMaskProcessor.downsample`` of
    diffusers==0.29.2 (third-party; pinned at requirements.txt:10; imported at
    src/models/base/mamba_layer.py:10; live call sites
    src/models/base/mamba_layer.py:1962 and :1973), and the index rule
    ``mask.view(-1).int().nonzero().view(-1)`` at mamba_layer.py:1963, 1974.
    diffusers is not under /root/reference and not installable here; the
    semantics are the published ones.  Pinned by the
    known-answer cases in tests/test_oracle.py (all-ones -> every token,
    rectangle [300:480, 180:400] of a 576x576 mask -> 594 / 154 / 36 tokens at
    L = 5184 / 1296 / 324).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

__all__ = ["downsample", "mask_to_index"]


def downsample(mask: torch.Tensor, batch_size: int, num_queries: int, value_embed_dim: int) -> torch.Tensor:
    """mask: (b, H, W) -> (batch_size, num_queries, value_embed_dim)."""
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
    """mask: (b, 1, H, W) as the pipeline passes it -> ascending int64 index list, exactly
    the expression used at mamba_layer.py:1962-1963 (values in (-1, 1) truncate to 0)."""
    m = downsample(mask[:, 0, :, :], mask.shape[0], num_queries, 1)
    return m.view(-1).int().nonzero().view(-1)
