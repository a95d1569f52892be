"""actalker_b200 — B200-native (sm_100a) implementation of ONE hot path of qazi0/ACTalker:
the masked selective-state-space control layer (reference src/models/base/mamba_layer.py:1902-1986).

    from actalker_b200 import SS2D_cond_v10, SS2D_Unit, selective_scan_fn, MAMBA_AVAILABLE

Python here is the host-side mirror of the reference's module / operator interface; the arithmetic is in
hand-written CUDA kernels behind the C-ABI declared in include/actalker_b200.h.  No CPU fallback.
"""
from .selective_scan_interface import MAMBA_AVAILABLE, a_kind_of, selective_scan_fn  # noqa: F401
from .mamba_layer import SS2D_Unit, SS2D_cond_v10, SS2D_cond_v10_wo_id, SS2D_cond_v8, SS2D_cond_v9  # noqa: F401
from .mask import MaskIndexCache, downsample, mask_to_index  # noqa: F401
from .host_api import HostStreamedLayer  # noqa: F401
from .sharded import ShardPlan, ShardedSS2DCondV10  # noqa: F401

__version__ = "0.1.0"
