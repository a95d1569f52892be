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

__all__ = ["ShardPlan", "ShardedSS2DCondV10", "BatchShardedCall", "all_gather_slices"]


@dataclass(frozen=True)
class ShardPlan:
    """Partition of B' (batch mode: balanced, remainder over the first ranks; block mode: equal blocks of ceil(B'/P) whose
    slots tile an all-gather buffer) or of d_inner (channel mode) over `world` ranks."""
    mode: str
    world: int
    extent: int          # B' or d_inner
    granule: int = 1     # channel mode: slices are multiples of 8 channels (16-byte rows), equal on every rank

    def __post_init__(self):
        if self.mode not in ("batch", "block", "channel"):
            raise ValueError("mode must be 'batch', 'block' or 'channel'")
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
        if self.mode == "block":      # equal blocks of ceil(extent / world); the last ones shorter or empty
            c = -(-self.extent // self.world)
            return min(rank * c, self.extent), min((rank + 1) * c, self.extent)
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


class BatchShardedCall:
    """ONE layer call split batch-first over the ranks ("assign (batch x CFG) blocks first"): every rank
    holds the call's full inputs (the denoiser around the layer is replicated), computes the layer for its block of whole
    frames and the result is all-gathered, so that every rank leaves with the full (B', L, d_model) output — the strong
    scaling of the live caller's B' = 4 x 25 = 100 call (pipeline ...two_ip.py:712).

    Partition: equal blocks of ceil(B'/P) frames, the last blocks shorter or empty.  The gather buffer has P * ceil(B'/P)
    frame slots and the result is its first B' frames, so the NCCL all-gather (equal contributions) writes the output in
    place: no padding copy, no re-layout.  Frames are independent, so the result equals the one-GPU call bit for bit as long
    as both take the same scan launch shape; a rank left with fewer than ~5 frames (B' = 25 over 4 or 8 GPUs) would take the
    two-level scan, which re-associates fp32 sums (differences of one 16-bit rounding step): `exact=True` pins the unsplit
    call's launch shape instead (a 4-frame block then scans in 0.84 instead of 0.75 ms).
    tiles > 1: a rank's block is cut into that many pieces; the gather of piece i runs on a side stream under the compute
    of piece i+1 (the pieces' slots are strided in the output, so each piece gathers into its own buffer and one copy per
    piece, also on the side stream, places it)."""

    MIN_FRAMES_PER_TILE = 25   # below ~1000 scan CTAs per launch the scan is latency-bound: cutting further costs more than
                               # the overlapped gather saves (B200, N=2, B'=25: 13 frames in two pieces 1.75 ms, in one 1.2 ms)

    def __init__(self, layer: SS2D_cond_v10, group=None, tiles: int = 1, gather: str = "nccl", exact: bool = False):
        """gather="nccl": NCCL all-gather of the blocks (in place in the padded output).
        gather="p2p": fused compute + collective — the out_proj kernel's epilogue stores every output tile straight into
        all ranks' gather buffers over NVLink peer memory (one TMA store per rank and tile; CUDA IPC buffers of
        PeerGatherBuffers), then a stream-ordered 4-byte all-reduce orders readers behind writers.  The returned tensor
        aliases this rank's buffer and stays valid until the call after next (buffers alternate)."""
        if gather not in ("nccl", "p2p"):
            raise ValueError("gather must be 'nccl' or 'p2p'")
        self.layer, self.group, self.max_tiles, self.gather = layer, group, max(1, int(tiles)), gather
        self.tiles = self.max_tiles
        self.exact = bool(exact)
        self._peer = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._side = None
        self._events = []

    def plan(self, Bp: int) -> "ShardPlan":
        return ShardPlan("block", self.world, Bp)

    def phase_ms(self):
        """Device time of the last call's phases on this rank: compute (all pieces) and the exposed tail after the last
        piece's compute (its gather).  Needs a synchronised device."""
        if not self._events:
            return None
        s, c, e = self._events
        return {"compute": s.elapsed_time(c), "gather_tail": c.elapsed_time(e)}

    def __call__(self, x, id_emb, conds, masks):
        if self.exact and self.world > 1:
            from . import mamba_layer as ml
            prev, ml.SCAN_BATCH_HINT = ml.SCAN_BATCH_HINT, x.shape[0]
            try:
                return self._call(x, id_emb, conds, masks)
            finally:
                ml.SCAN_BATCH_HINT = prev
        return self._call(x, id_emb, conds, masks)

    def _call(self, x, id_emb, conds, masks):
        layer, P, rank = self.layer, self.world, self.rank
        Bp, L, dm = x.shape
        if P == 1:
            return layer(x, id_emb, conds, masks)
        cuda = x.is_cuda                                            # the gloo tests drive the same code on CPU tensors
        chunk = -(-Bp // P)
        self.tiles = max(1, min(self.max_tiles, chunk // self.MIN_FRAMES_PER_TILE)) if cuda else self.max_tiles
        lo, hi = self.plan(Bp).bounds(rank)
        out = torch.empty((P * chunk, L, dm), dtype=x.dtype, device=x.device)
        main = torch.cuda.current_stream(x.device) if cuda else None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if cuda else None

        def compute(a, b, dst):
            """the layer for frames [a, b), written straight into dst (out_proj stores there: no copy)"""
            if b > a:
                y = layer(x[a:b], id_emb[a:b], conds[a:b], masks, out=dst[:b - a]) if cuda else layer(x[a:b], id_emb[a:b], conds[a:b], masks)
                if y.data_ptr() != dst.data_ptr():
                    dst[:b - a] = y

        if cuda:
            ev[0].record(main)
        if self.gather == "p2p" and cuda:
            self.tiles = 1
            numel = P * chunk * L * dm
            if self._peer is None or self._peer.numel != numel or self._peer.dtype != x.dtype:
                if self._peer is not None:
                    self._peer.close()
                self._peer = PeerGatherBuffers(P, rank, numel, x.dtype, x.device, self.group)
            ptrs, mine = self._peer.next()
            off = rank * chunk * L * dm * x.element_size()          # this rank's block, same offset in every buffer
            if hi > lo:
                layer(x[lo:hi], id_emb[lo:hi], conds[lo:hi], masks, out_peers=[p + off for p in ptrs])
            ev[1].record(main)
            with _timed("all_gather", x.device):
                self._peer.fence()
            ev[2].record(main)
            self._events = ev
            return _device_view(mine, (P * chunk, L, dm), x.dtype, x.device)[:Bp]
        if self.tiles == 1:
            mine = out[rank * chunk:(rank + 1) * chunk]
            compute(lo, hi, mine)
            if cuda:
                ev[1].record(main)
            with _timed("all_gather", x.device):       # in place: the rank's block already sits in its slot of the buffer
                dist.all_gather_into_tensor(out.view(-1), mine.reshape(-1), group=self.group)
        else:
            if cuda and self._side is None:
                self._side = torch.cuda.Stream(x.device)
            side = self._side
            tr = -(-chunk // self.tiles)                              # frames per piece (the last may be shorter / empty)
            keep = []
            for t in range(self.tiles):
                a, b = min(lo + t * tr, hi), min(lo + (t + 1) * tr, hi)
                mine = torch.empty((tr, L, dm), dtype=x.dtype, device=x.device)
                compute(a, b, mine)
                gathered = torch.empty((P, tr, L, dm), dtype=x.dtype, device=x.device)
                n = min(tr, chunk - t * tr)
                if cuda:
                    done = torch.cuda.Event()
                    done.record(main)
                    side.wait_event(done)
                    with torch.cuda.stream(side):
                        dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1), group=self.group)
                        if n > 0:
                            out.view(P, chunk, L, dm)[:, t * tr:t * tr + n] = gathered[:, :n]
                else:
                    dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1), group=self.group)
                    if n > 0:
                        out.view(P, chunk, L, dm)[:, t * tr:t * tr + n] = gathered[:, :n]
                keep.append((mine, gathered))
            if cuda:
                ev[1].record(main)
                main.wait_stream(side)
        if cuda:
            ev[2].record(main)
            self._events = ev
        return out[:Bp]


def _device_view(ptr: int, shape, dtype, device) -> torch.Tensor:
    """A torch tensor over memory this library allocated (a peer gather buffer): the CUDA array interface with a 16-bit
    integer type string, reinterpreted as the activation dtype (bf16 has no numpy type string)."""
    numel = 1
    for n in shape:
        numel *= n
    es = torch.empty(0, dtype=dtype).element_size()

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (numel,), "typestr": {2: "<i2", 4: "<i4"}[es], "data": (int(ptr), False), "version": 2}
    with torch.cuda.device(device):
        t = torch.as_tensor(h, device=device)
    return t.view(dtype).view(shape)


class PeerGatherBuffers:
    """Gather buffers for the fused push all-gather (gather="p2p"): every rank owns `depth` buffers of
    (world, rows, Ds) elements (plain cudaMalloc, C-ABI actk_peer_buffer_*) and maps every other rank's buffers into
    its own device's address space through CUDA IPC, so the merge kernel can store its slice into all of them over
    NVLink.

    Ordering: call i writes buffer i % depth on every rank, then a stream-ordered 4-byte all-reduce separates the
    writers from the readers (a rank's reduction kernel starts after its own merge kernel, and completes only once
    every rank has joined).  With depth 2 a buffer is rewritten two calls later, after every rank has passed the
    barrier of the call in between, which it enqueues behind its reads of this buffer."""

    def __init__(self, world: int, rank: int, numel: int, dtype, device, group=None, depth: int = 2):
        lib = _lib.load()
        self.world, self.rank, self.group, self.depth, self.calls = world, rank, group, depth, 0
        self.numel, self.dtype, self.device = numel, dtype, device
        nbytes = numel * torch.empty(0, dtype=dtype).element_size()
        self.local, handles = [], []
        with torch.cuda.device(device):
            for _ in range(depth):
                p = ct.c_void_p()
                _lib.check(lib.actk_peer_buffer_alloc(nbytes, ct.byref(p)), "actk_peer_buffer_alloc")
                h = ct.create_string_buffer(64)
                _lib.check(lib.actk_peer_buffer_export(p, h), "actk_peer_buffer_export")
                self.local.append(p.value)
                handles.append(h.raw)
            everyone = [None] * world
            dist.all_gather_object(everyone, handles, group=group)
            self._opened = []
            self.ptrs = [[0] * world for _ in range(depth)]
            for r, hs in enumerate(everyone):
                for d in range(depth):
                    if r == rank:
                        self.ptrs[d][r] = self.local[d]
                    else:
                        p = ct.c_void_p()
                        _lib.check(lib.actk_peer_buffer_open(hs[d], ct.byref(p)), "actk_peer_buffer_open")
                        self._opened.append(p.value)
                        self.ptrs[d][r] = p.value
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(group=group)

    def next(self):
        d = self.calls % self.depth
        self.calls += 1
        return self.ptrs[d], self.local[d]

    def fence(self):
        dist.all_reduce(self._flag, group=self.group)    # stream-ordered: no host synchronisation

    def close(self):
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.actk_peer_buffer_close(ct.c_void_p(p))
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for p in self.local:
                lib.actk_peer_buffer_free(ct.c_void_p(p))
        self._opened, self.local = [], []


class ShardedSS2DCondV10(torch.nn.Module):
    """Wraps a (replicated) SS2D_cond_v10.  forward takes the same arguments as the layer.
    batch mode: each rank passes ITS rows of B' (the caller shards the batch) and gets its rows back.
    channel mode: every rank passes the same full inputs and gets the full output."""

    def __init__(self, layer: SS2D_cond_v10, mode: str = "channel", group=None, gather: str = "nccl"):
        """gather="nccl": one NCCL all-gather of the merged slices (the exchange BASELINE.json names).
        gather="p2p": the merge kernel itself stores every slice into all ranks' gather buffers over NVLink peer
        memory (fused compute + collective), followed by a 4-byte barrier all-reduce."""
        super().__init__()
        if gather not in ("nccl", "p2p"):
            raise ValueError("gather must be 'nccl' or 'p2p'")
        self.layer, self.mode, self.group, self.gather = layer, mode, group, gather
        self._peer = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.plan = ShardPlan("channel", self.world, layer.d_inner, 8) if mode == "channel" else None

    def forward(self, x, id_emb, conds, masks):
        layer = self.layer
        if self.mode == "batch" or self.world == 1:
            return layer(x, id_emb, conds, masks)
        lo, hi = self.plan.bounds(self.rank)
        xz1, xz2, tail1, tail2, m1, m2 = layer.project_inputs(x, id_emb, conds, masks)
        if self.gather == "p2p":
            Bp, L, Ds = xz1.shape[0], xz1.shape[1], hi - lo
            numel = self.world * Bp * L * Ds
            if self._peer is None or self._peer.numel != numel or self._peer.dtype != xz1.dtype:
                if self._peer is not None:
                    self._peer.close()
                self._peer = PeerGatherBuffers(self.world, self.rank, numel, xz1.dtype, x.device, self.group)
            ptrs, mine = self._peer.next()
            layer.scan_core(xz1, xz2, tail1, tail2, m1, m2, ch_slice=(lo, hi), push=(ptrs, self.rank))
            with _timed("all_gather", x.device):
                self._peer.fence()
            y = self.gathered_layernorm(mine, (self.world, Bp, L, Ds), xz1.dtype, x.device)
            return layer._out_proj(y)
        else:
            merged = layer.scan_core(xz1, xz2, tail1, tail2, m1, m2, ch_slice=(lo, hi))  # (B', L, Ds)
            with _timed("all_gather", x.device):
                gathered = all_gather_slices(merged, self.group)                           # (P, B', L, Ds)
        return layer._out_proj(self.gathered_layernorm(gathered))

    def gathered_layernorm(self, gathered, shape=None, dtype=None, device=None) -> torch.Tensor:
        """LayerNorm over the gathered (rank, row, slice) layout: `gathered` is a (P, B', L, Ds) tensor, or the raw
        device address of a peer gather buffer together with shape / dtype / device."""
        if isinstance(gathered, torch.Tensor):
            keep, shape, dtype, device = gathered, tuple(gathered.shape), gathered.dtype, gathered.device
            gathered_ptr = keep.data_ptr()
        else:
            gathered_ptr = int(gathered)
        lib = _lib.load()
        P, Bp, L, Ds = shape
        norm = self.layer.out_norm
        out = torch.empty((Bp, L, P * Ds), dtype=dtype, device=device)
        gamma, beta = norm.weight.to(dtype), norm.bias.to(dtype)
        with torch.cuda.device(device), _timed("gathered_ln", device):
            _lib.check(lib.actk_gathered_layernorm_fwd(ct.c_void_p(gathered_ptr), P, Bp * L, Ds, _ptr(gamma), _ptr(beta),
                                                       float(norm.eps), _ptr(out), _DTYPES[dtype],
                                                       ct.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
                       "actk_gathered_layernorm_fwd")
        return out
