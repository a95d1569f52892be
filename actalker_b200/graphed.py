"""CUDA-graph capture of the layer.

Once the mask indices and the derived weights are cached, a layer call has static shapes and performs no host
synchronisation, so the whole forward — cuBLAS GEMMs, the tensor maps baked into the scan launch, the chain-mode
memset, the merge/LayerNorm kernel — can be captured once and replayed.  For the launch-bound small-batch calls
(single frames, the 18x18 / 36x36 resolutions of the UNet) this removes ~25 launches' worth of CPU time per call.

    g = GraphedLayer(layer, x, id_emb, conds, masks)      # warms up, then captures
    y = g(x2, id2, conds2)                                # copies into the static inputs, replays, returns a view
"""
from __future__ import annotations

import torch

__all__ = ["GraphedLayer"]


class GraphedLayer:
    def __init__(self, layer, x, id_emb, conds, masks, warmup: int = 3):
        self.layer, self.masks = layer, masks
        self.x, self.id_emb, self.conds = x.clone(), id_emb.clone(), conds.clone()
        side = torch.cuda.Stream(x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(warmup):          # fills the mask-index / derived-weight caches (their one-off host syncs)
                layer(self.x, self.id_emb, self.conds, masks)
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.y = layer(self.x, self.id_emb, self.conds, masks)

    def __call__(self, x, id_emb, conds):
        self.x.copy_(x)
        self.id_emb.copy_(id_emb)
        self.conds.copy_(conds)
        self.graph.replay()
        return self.y
