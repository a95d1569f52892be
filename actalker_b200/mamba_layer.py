"""Drop-in `SS2D_cond_v10` / `SS2D_Unit` running on the sm_100a kernels.

Mirror of the reference's live masked Mamba control layer, src/models/base/mamba_layer.py:
    SS2D_Unit      :1394-1553   (bidirectional K=2 selective-scan unit)
    SS2D_cond_v10  :1902-1986   (audio + expression branches, mask gather/scatter, LayerNorm, out_proj)
Same constructor arguments, same `forward(x, id_emb, conds, masks)`, same parameter names, shapes and dtypes
(SURVEY.md Appendix C), so `unet.load_state_dict(strict=True)` (Inference.py:124-127) and the fp32 re-cast of
`A_logs | Ds | dt_projs_bias` (Inference.py:430-433) keep working unchanged.

What runs where (16-bit activations — the reference's inference dtype; everything is this repo's CUDA, C-ABI
include/actalker_b200.h):
        actk_gemm_tn_fwd           in_proj1/2, id/audio/exp projections (+SiLU), x_proj, dt_proj, out_proj: persistent
                                   TMA + tcgen05 kernel, same contractions and rounding points as the reference's
                                   nn.Linear / einsum calls                                    (:1960-1961, :1521-1523, :1985)
        actk_masked_scan_fwd       gather -> tail concat -> both scan directions -> scatter   (:1963-1970, :1505-1548)
        actk_merge_layernorm_fwd   direction sum, passthrough rows, branch sum, out_norm      (:1542-1547, :1983-1984)
    fp32 activations: the projections stay fp32 GEMMs in torch (tensor cores would round the operands; the reference's own
    arithmetic for that dtype); the scan and merge kernels are the same.  ACTK_TC_GEMM=0 forces the torch GEMMs.
The mask -> index computation is the reference's own expression, cached per mask (mask.py).
Forward only: inference runs under no_grad (pipeline ...two_ip.py:351); backward is not built.
"""
from __future__ import annotations

import ctypes as ct
import math
import os
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import gemm
from .mask import MaskIndex, MaskIndexCache
from .selective_scan_interface import _DTYPES, _ptr, _stream, a_kind_of

__all__ = ["SS2D_Unit", "SS2D_cond_v10", "SS2D_cond_v10_wo_id", "SS2D_cond_v8", "SS2D_cond_v9", "MAMBA_AVAILABLE"]

try:
    _lib.load()
    MAMBA_AVAILABLE = True
except _lib.LibraryMissing:
    MAMBA_AVAILABLE = False

_N = 16  # d_state compiled into the kernels

# Optional kernel timing hook (bench.py): when set to a dict, every C-ABI launch is counted (TIMING["launches"]) and
# bracketed by CUDA events on the launching stream, the (name, start, end) triples appended to TIMING["events"];
# TIMING["only"] = {names} restricts the bracketing to those launches.
TIMING = None


class _timed:
    def __init__(self, name, device):
        self.name, self.device = name, device

    def __enter__(self):
        self.on = False
        if TIMING is not None:
            TIMING["launches"] = TIMING.get("launches", 0) + 1
            only = TIMING.get("only")        # optional set of names: bracket just those launches (each pair of event
            self.on = only is None or self.name in only   # records costs the stream ~5 us; 14 of them are 3 % of a 2 ms step)
        if self.on:
            self.s, self.e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if self.on:
            self.e.record(torch.cuda.current_stream(self.device))
            TIMING.setdefault("events", []).append((self.name, self.s, self.e))
        return False


# merge + LayerNorm + out_proj as one tcgen05 kernel where the shape is built (D = 640, 16-bit).
# Off by default (ACTK_FUSE_LN_OUT=1 turns it on): measured on B200 at config 2 it takes 0.46 ms against 0.16 + 0.04 ms
# for the merge kernel + cuBLAS GEMM — one CTA per SM around a 160 KB operand tile cannot keep enough HBM loads in
# flight while it normalises rows.
FUSE_LN_OUT_PROJ = os.environ.get("ACTK_FUSE_LN_OUT", "0") == "1"
BATCH_IN_PROJ = os.environ.get("ACTK_BATCH_IN_PROJ", "1") != "0"   # one batched GEMM for in_proj1 / in_proj2
_SIDE_STREAMS = {}
SIDE_STREAM = os.environ.get("ACTK_SIDE_STREAM", "1") != "0"
# 16-bit activations: the dense projections run on this repo's tcgen05 kernel (gemm.py); "0" routes them to torch.
TC_GEMM = os.environ.get("ACTK_TC_GEMM", "1") != "0"


class _Fork:
    """Run the tiny id / condition-token GEMMs (a dozen ~4 us launches, ~55 us per call when serialised) on a side
    stream beside the large in_proj / x_proj GEMMs of the latent tokens.  Fork: the side stream waits for the current
    one; join: the current stream waits for the side stream — the pattern CUDA-graph capture accepts."""

    def __init__(self, device):
        self.main = torch.cuda.current_stream(device)
        if SIDE_STREAM:
            # one side stream per (device, calling stream): a fork only waits for ITS main stream, so sharing a side
            # stream between callers on different streams would let one call reuse side-pool blocks another still reads
            key = (device.index if device.index is not None else torch.cuda.current_device(), self.main.cuda_stream)
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
            self.side = _SIDE_STREAMS[key]
        else:
            self.side = None
        self.wait = True

    def __call__(self, wait: bool):
        """`with fork(wait=False)`: re-enter the side stream WITHOUT waiting for what the current stream queued since
        the first entry (only correct when the side work depends on nothing issued there in between)."""
        self.wait = wait
        return self

    def __enter__(self):
        if self.side is not None:
            if self.wait:
                self.side.wait_stream(self.main)
            self.wait = True
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.side is not None:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *tensors):
        # No record_stream on the results: it defers the allocator's reuse of those blocks behind event queries and
        # produced sporadic cudaMalloc stalls (~1 ms) in the layer loop.  It is not needed here: side-stream blocks are
        # only ever re-used by side-stream work, and every fork starts by waiting for everything the current stream
        # has queued — including the kernels that read these results.
        if self.side is not None:
            self.main.wait_stream(self.side)


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


_FUSED_KSLABS = (2, 3, 5)   # 16-wide rank slabs the fused dt_proj kernels are built for (dt_rank_pad 32 / 48 / 80)
# SURVEY §8 row f1.  True (or ACTK_FUSE_DT=1 in the environment): with 16-bit activations the scan kernel computes
# dt_proj itself per tile (tcgen05.mma into tensor memory) and no delta tensor exists.  Off by default: measured on
# B200 at config 2 the per-tile issue / wait instructions cost the issue-bound scan +0.19 ms (1.51 -> 1.70 ms) while
# the two cuBLAS dt_proj GEMMs they replace cost 0.14 ms.
FUSE_DT_PROJ = os.environ.get("ACTK_FUSE_DT", "0") == "1"
# ACTK_LEAN_SCAN=1: with 16-bit activations under all-ones masks (the shipped pipeline, Inference.py:545-546) the x_proj
# launch also writes the B|C columns as fp32 and the scan takes the lean kernel (csrc/masked_scan_lean.cu: warp-autonomous
# tiles, a third less per-tile code).  Bit-identical results; OFF by default: measured on B200 at config 2 it runs in 1.46 ms
# against 1.43 ms — the 8-step loops, not the code around them, bound the scan (profiles/r02_lean_scan.txt).
LEAN_SCAN = os.environ.get("ACTK_LEAN_SCAN", "0") == "1"


def _rank_pad(R: int):
    """(padded rank, fusable): the dt_proj input of each direction occupies a block of `padded rank` x_dbl columns.
    Ranks up to 80 (d_model <= 1280 with dt_rank='auto') round up to a width the in-kernel tcgen05 dt_proj handles."""
    for ks in _FUSED_KSLABS:
        if R <= 16 * ks:
            return 16 * ks, True
    return _pad8(R), False


class SS2D_Unit(nn.Module):
    """Parameters and initialisers of mamba_layer.py:1394-1502; forward == forward_core (:1505-1548)."""

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
            [nn.Linear(self.d_inner, c, bias=False, **fk).weight for _ in range(K)], dim=0))        # (K, R+2N, D)
        dt_projs = [self.dt_init(self.dt_rank, self.d_inner, dt_scale, dt_init, dt_min, dt_max, dt_init_floor, **fk)
                    for _ in range(K)]
        self.dt_projs_weight = nn.Parameter(torch.stack([p.weight for p in dt_projs], dim=0))       # (K, D, R)
        self.dt_projs_bias = nn.Parameter(torch.stack([p.bias for p in dt_projs], dim=0))           # (K, D)
        self.A_logs = self.A_log_init(d_state, self.d_inner, copies=K, device=device)               # (K*D, N) fp32
        self.Ds = self.D_init(self.d_inner, copies=K, device=device)                                # (K*D,)   fp32
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else None   # constructed, never applied (as upstream)
        self._derived = None
        self._derived_key = None

    # -- initialisers: the synthetic parameter distribution of the benchmarks (mamba_layer.py:1450-1502)
    @staticmethod
    def dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4, **fk):
        proj = nn.Linear(dt_rank, d_inner, bias=True, **fk)
        std = dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(d_inner, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = dt.clamp(min=dt_init_floor)
        with torch.no_grad():
            proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))     # softplus^-1(dt)
        return proj

    @staticmethod
    def A_log_init(d_state, d_inner, copies=1, device=None, merge=True):
        a = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(d_inner, 1)
        if copies > 1:
            a = a.repeat(copies, 1) if merge else a[None].repeat(copies, 1, 1)
        return nn.Parameter(a.contiguous())

    @staticmethod
    def D_init(d_inner, copies=1, device=None, merge=True):
        d = torch.ones(d_inner * (copies if copies > 1 and merge else 1), dtype=torch.float32, device=device)
        if copies > 1 and not merge:
            d = d.view(1, -1).repeat(copies, 1)
        return nn.Parameter(d)

    # -- derived, kernel-friendly weights; rebuilt when a parameter is replaced or modified in place
    def derived(self):
        ps = (self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)
        key = tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in ps)
        if self._derived is not None and key == self._derived_key:
            return self._derived
        K, D, R, N = self.num_direction, self.d_inner, self.dt_rank, self.d_state
        if K != 2 or N != _N:
            raise NotImplementedError(f"kernels are built for num_direction=2, d_state={_N} (the live layer); "
                                      f"got K={K}, d_state={N}")
        with torch.no_grad():
            Rp, fusable = _rank_pad(R)
            xw = 2 * K * N + K * Rp
            wx = self.x_proj_weight                                   # (K, R+2N, D)
            w = wx.new_zeros(xw, D)                                   # rows [B_0|C_0|B_1|C_1 | dt_0 (Rp) | dt_1 (Rp)]
            for k in range(K):
                w[k * 2 * N:k * 2 * N + 2 * N] = wx[k, R:R + 2 * N]   # [B_k | C_k]
                w[2 * K * N + k * Rp:2 * K * N + k * Rp + R] = wx[k, :R]   # rank rows, zero-padded to Rp
            wbd = wx.new_zeros(K * Rp, K * D)                         # block-diagonal dt_proj for the GEMM route
            for k in range(K):
                wbd[k * Rp:k * Rp + R, k * D:(k + 1) * D] = self.dt_projs_weight[k].t()
            A = (-torch.exp(self.A_logs.float())).contiguous()        # (K*D, N), as mamba_layer.py:1530
            d = {"xw": xw, "w_xproj": w.contiguous(), "w_dt": wbd.contiguous(),
                 "rank_pad": Rp, "fusable": fusable, "A": A,
                 "Ds": self.Ds.float().contiguous().view(-1), "dt_bias": self.dt_projs_bias.float().contiguous().view(-1),
                 "a_kind": a_kind_of(A) if A.is_cuda else _lib.ACTK_A_GENERAL}
        self._derived, self._derived_key = d, key
        return d

    def weights_for(self, dtype, lo: int = 0, hi: int = None):
        """Per-(activation dtype, channel slice) views of the derived weights, cached beside them (rebuilt when a
        parameter changes): no cast, slice or gather kernel runs per call, so nothing the side stream consumes is ever
        produced by fresh main-stream work.
          w_xproj (xw, D)       x_proj rows [B_0|C_0|B_1|C_1|dt_0|dt_1]
          w_dt    (2Rp, 2Dk)    block-diagonal dt_proj, torch.mm form (fp32 route)
          w_dt_nk (2Dk, 2Rp)    the same as an nn.Linear weight (tensor-core route)
          A (2Dk, N), Ds, dt_bias (2Dk) fp32"""
        dv = self.derived()
        D = self.d_inner
        hi = D if hi is None else hi
        key = ("weights", dtype, lo, hi)
        if key not in dv:
            with torch.no_grad():
                w_dt, A, Ds, dtb = dv["w_dt"], dv["A"], dv["Ds"], dv["dt_bias"]
                if (lo, hi) != (0, D):   # columns / rows [k*D + lo, k*D + hi) of both directions
                    cols = torch.cat([torch.arange(k * D + lo, k * D + hi, device=A.device) for k in range(2)])
                    w_dt, A, Ds, dtb = w_dt[:, cols], A[cols].contiguous(), Ds[cols].contiguous(), dtb[cols].contiguous()
                dv[key] = {"w_xproj": dv["w_xproj"].to(dtype).contiguous(), "w_dt": w_dt.to(dtype).contiguous(),
                           "w_dt_nk": w_dt.t().to(dtype).contiguous(), "A": A, "Ds": Ds, "dt_bias": dtb}
        return dv[key]

    def dt_image(self, lo: int, hi: int, dtype) -> torch.Tensor:
        """Tensor-core operand image of dt_projs_weight[:, lo:hi] for the fused dt_proj (actk_pack_dt_proj_weight),
        cached with the other derived weights (rebuilt when a parameter changes)."""
        dv = self.derived()
        key = ("dt_image", lo, hi, dtype)
        if key not in dv:
            lib = _lib.load()
            w = self.dt_projs_weight.detach()[:, lo:hi].to(dtype).contiguous()          # (K, Dk, R)
            Dk, R, rp = hi - lo, self.dt_rank, dv["rank_pad"]
            img = torch.empty(lib.actk_dt_proj_image_bytes(Dk, rp, w.element_size()), dtype=torch.uint8, device=w.device)
            with torch.cuda.device(w.device):
                _lib.check(lib.actk_pack_dt_proj_weight(_ptr(w), Dk, R, rp, _DTYPES[dtype], _ptr(img), _stream(w)),
                           "actk_pack_dt_proj_weight")
            dv[key] = img
        return dv[key]

    def forward_core(self, x: torch.Tensor):
        """x: (B, D, L) -> (B, D, L).  Same contract as upstream; internally token-major."""
        Bsz, D, L = x.shape
        xt = x.permute(0, 2, 1)
        if not xt.is_contiguous():
            xt = xt.contiguous()
        idx = torch.arange(L, dtype=torch.int32, device=x.device)
        ydir = _scan_branches([self], [xt], [None], [idx], [L], Bsz, L)[0][0]   # (2, B, L, D)
        return (ydir[0] + ydir[1]).permute(0, 2, 1)

    def forward(self, input):
        return self.forward_core(input)


SCAN_SEGMENTS = None   # None: choose per call (auto_segments); an int forces that many chunks (tests, tuning)
SCAN_BATCH_HINT = None # an int: choose the two-level launch shape as a call of that many frames would (a batch-split call that
                       # must reproduce the unsplit call bit for bit: the two-level scan re-associates fp32 sums)
SCAN_CHAIN = None      # None: choose per call (auto_chain); an int forces that many chained chunks (0/1 = off)
POISON_OUTPUTS = False # tests: pre-fill the scan output with NaN so a row the kernel fails to write cannot go unnoticed


def auto_chain(n_ctas: int, min_tiles: int, n_sms: int = 148) -> int:
    """Chained chunks per sequence (load balancing, masked_scan.cu MODE 2): worthwhile once the launch fills the
    GPU (>= 4 CTAs per SM) and sequences are long enough to cut into chunks of ~20 tiles (measured on B200 at
    config 2, scan ms for 0/4/8/16/24/40 chunks: 1.79/1.60/1.51/1.49/1.50/1.53 general, 1.52/1.31/1.23/1.21/1.22/1.26
    power)."""
    if n_ctas < 4 * n_sms or min_tiles < 80:
        return 0
    return min(min_tiles // 20, 64)



def auto_segments(n_ctas: int, min_tiles: int, n_sms: int = 148) -> int:
    """Chunks per sequence for the two-level scan.  The extra state-only pass costs ~75 % more work, so cutting time
    only pays when the independent (batch x branch x direction x channel-block) axes leave warp schedulers EMPTY:
    about one 2-warp CTA per SM or fewer (single-frame calls, BASELINE config 5's one long sequence).  From one CTA per
    SM upwards a single level is faster even at low occupancy — measured on B200 for the channel-sharded slices of
    config 2: 500 CTAs 1.02 ms single level vs 1.74 ms in two segments, 300 CTAs 1.01 vs 1.16, 200 CTAs 0.70 vs 0.90
    (tools/bench_sliced.py).  When it does cut, it aims at about 6 CTAs per SM with chunks of >= 8 tiles."""
    if n_ctas >= 5 * n_sms // 4 or min_tiles < 16:    # 160 CTAs (B'=4 at 72x72): two levels 0.75 vs 0.84 ms per call
        return 1
    return max(1, min(-(-6 * n_sms // n_ctas), min_tiles // 8))


def _gather_rows(t: torch.Tensor, idx32: torch.Tensor) -> torch.Tensor:
    """t[:, idx, :] for a contiguous (B, rows, row) CUDA tensor and an int32 index list on its device: the own row-gather
    kernel (16-byte vectors, one warp per row) where rows are multiples of 16 bytes, torch's index_select otherwise."""
    Bt, rows, width = t.shape
    if not (t.is_cuda and t.is_contiguous() and (width * t.element_size()) % 16 == 0 and t.data_ptr() % 16 == 0):
        return t.index_select(1, idx32.long())
    out = torch.empty((Bt, idx32.numel(), width), dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device), _timed("gather_rows", t.device):
        _lib.check(_lib.load().actk_gather_rows(_ptr(t), _ptr(idx32), _ptr(out), Bt, rows, idx32.numel(),
                                                width * t.element_size(), _stream(t)), "actk_gather_rows")
    return out


def _scan_branches(units: List[SS2D_Unit], xzs, tails, idxs, n_sels, Bp: int, L: int, idx64s=None, ch_slice=None):
    """Shared launcher: x_proj / dt_proj of every branch + one actk_masked_scan_fwd for all given branches.
    xzs[i]: (Bp, L, D) contiguous; tails[i]: (Bp, n_tail, D) or None; idxs[i]: int32 (n_sel,).
    ch_slice=(lo, hi): scan only channels [lo, hi) of every direction (multi-GPU channel sharding): x_proj still
    contracts over all D channels (B|C are replicated), delta / A / D / dt_bias / u are sliced.
    Returns per-branch (ydir (2, Bp, L, Dk), xz_k (Bp, L, Dk)) with Dk = hi - lo (D when unsliced); rows of
    unselected tokens of ydir are uninitialised.

    16-bit activations: the projections of ALL branches (latent and tail tokens) are two launches of this repo's
    tensor-core kernel (gemm.py) — x_proj, then dt_proj reading the dt columns of x_dbl in place.  fp32: torch GEMMs,
    the tail tokens' small ones on a side stream."""
    lib = _lib.load()
    x0 = xzs[0]
    if not x0.is_cuda:
        raise RuntimeError("actalker_b200 layers run on CUDA tensors only (no CPU path)")
    if x0.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported activation dtype {x0.dtype}")
    D = x0.shape[-1]
    lo, hi = (0, D) if ch_slice is None else ch_slice
    Dk = hi - lo
    sliced = Dk != D
    tc = TC_GEMM and gemm.usable(*xzs, *tails) and D % 8 == 0
    args = _lib.MaskedScanArgs()
    args.n_branches, args.Bp, args.L, args.D, args.N = len(units), Bp, L, Dk, _N
    args.dtype = _DTYPES[x0.dtype]
    keep, outs, jobs = [], [], []
    fork = None if tc else _Fork(x0.device)
    xproj = []                              # tensor-core route: x_proj problems of every branch, one launch
    for i, unit in enumerate(units):
        dv = unit.derived()
        xz, tail, n_sel = xzs[i], tails[i], n_sels[i]
        n_tail = 0 if tail is None else tail.shape[1]
        ydir = torch.empty((2, Bp, L, Dk), dtype=xz.dtype, device=xz.device)
        if POISON_OUTPUTS:
            ydir.fill_(float("nan"))
        xz_k = xz[..., lo:hi].contiguous() if sliced else xz
        outs.append((ydir, xz_k))
        b = args.br[i]
        b.n_sel, b.n_tail, b.a_kind = n_sel, n_tail, dv["a_kind"]
        if n_sel == 0:
            continue
        xw = dv["xw"]
        w = unit.weights_for(xz.dtype, lo, hi)
        w_x = w["w_xproj"]
        fused = FUSE_DT_PROJ and dv["fusable"] and xz.element_size() == 2
        # x_proj of the SELECTED tokens only, in sequence order (row p <-> latent token idx[p]).  Under a partial mask
        # the rows are gathered first when few are selected (the GEMM shrinks with them), else projected in place and
        # the narrow x_dbl rows gathered afterwards.
        sel64 = None if n_sel == L else (idx64s[i] if idx64s is not None else idxs[i].long())
        gather_after = sel64 is not None and 2 * n_sel > L
        src = xz if (sel64 is None or gather_after) else _gather_rows(xz, idxs[i])
        job = {"i": i, "unit": unit, "w": w, "xw": xw, "fused": fused, "n_sel": n_sel, "n_tail": n_tail, "tail": tail,
               "sel64": sel64 if gather_after else None, "idx32": idxs[i], "ydir": ydir, "xz_k": xz_k, "src": src}
        if tc:
            # lean scan kernel: the B|C columns also leave the x_proj launch as fp32 (Bp, n, 4N)
            lean = LEAN_SCAN and not fused and n_sel == L and Dk % 64 == 0
            job["bc32"] = torch.empty((Bp, n_sel, 4 * _N), dtype=torch.float32, device=xz.device) if lean else None
            job["xdbl"] = torch.empty((Bp, src.shape[1], xw), dtype=xz.dtype, device=xz.device)
            xproj.append(gemm.Problem(src.view(-1, D), w_x, job["xdbl"].view(-1, xw),
                                      f32=job["bc32"].view(-1, 4 * _N) if lean else None))
            job["bc32_tail"] = None
            if n_tail:
                job["xdbl_tail"] = torch.empty((Bp, n_tail, xw), dtype=xz.dtype, device=xz.device)
                if lean:
                    job["bc32_tail"] = torch.empty((Bp, n_tail, 4 * _N), dtype=torch.float32, device=xz.device)
                xproj.append(gemm.Problem(tail.view(-1, D), w_x, job["xdbl_tail"].view(-1, xw),
                                          f32=job["bc32_tail"].view(-1, 4 * _N) if lean else None))
            else:
                job["xdbl_tail"] = None
        else:
            with fork:                                                         # tail tokens: side stream
                job["xdbl_tail"] = F.linear(tail, w_x) if n_tail else None     # (Bp, n_tail, xw)
            job["xdbl"] = F.linear(src, w_x)
        jobs.append(job)
    if tc:
        gemm.run(xproj, name="gemm_xproj")
    dtproj = []
    for job in jobs:
        i, w, xw, n_sel, n_tail, tail = job["i"], job["w"], job["xw"], job["n_sel"], job["n_tail"], job["tail"]
        xdbl, xdbl_tail, fused = job["xdbl"], job["xdbl_tail"], job["fused"]
        if job["sel64"] is not None:
            xdbl = _gather_rows(xdbl, job["idx32"])                            # (Bp, n_sel, xw), sequence order
        if sliced and tail is not None:
            tail = tail[..., lo:hi].contiguous()
        delta = delta_tail = w_img = None
        if fused:
            # the scan kernel multiplies each 16-token tile of dt columns by w_dt on the tensor cores itself
            # (tcgen05.mma, fp32 accumulate, one rounding to the activation dtype) — no delta tensor, no dt_proj launch
            w_img = job["unit"].dt_image(lo, hi, xdbl.dtype)
        elif tc:
            # dt_proj of both directions as one product with the block-diagonal weight; the dt columns of x_dbl are
            # read in place (row pitch xw) — no copy of the (Bp*n_sel, 2Rp) slice, and the tail rows are one more
            # problem of the same launch instead of a concatenation.  Rows are in sequence order.
            delta = torch.empty((Bp, n_sel, 2 * Dk), dtype=xdbl.dtype, device=xdbl.device)
            dtproj.append(gemm.Problem(xdbl.view(-1, xw)[:, 4 * _N:], w["w_dt_nk"], delta.view(-1, 2 * Dk)))
            if n_tail:
                delta_tail = torch.empty((Bp, n_tail, 2 * Dk), dtype=xdbl.dtype, device=xdbl.device)
                dtproj.append(gemm.Problem(xdbl_tail.view(-1, xw)[:, 4 * _N:], w["w_dt_nk"], delta_tail.view(-1, 2 * Dk)))
        else:
            with fork(wait=False):    # depends on side-stream work and cached weights only
                delta_tail = torch.mm(xdbl_tail.view(Bp * n_tail, xw)[:, 4 * _N:], w["w_dt"]).view(Bp, n_tail, 2 * Dk) if n_tail else None
            delta = torch.mm(xdbl.view(Bp * n_sel, xw)[:, 4 * _N:], w["w_dt"]).view(Bp, n_sel, 2 * Dk)
        rp = job["unit"].derived()["rank_pad"] if fused else 0
        if args.xw not in (0, xw) or (args.xw != 0 and args.dt_rank_pad != rp):
            raise RuntimeError("branches disagree on the x_proj width / dt rank")
        args.xw, args.dt_rank_pad = xw, rp
        b = args.br[i]
        b.xz, b.tail, b.xdbl, b.xdbl_tail, b.delta = _ptr(job["xz_k"]), _ptr(tail), _ptr(xdbl), _ptr(xdbl_tail), _ptr(delta)
        b.delta_tail = _ptr(delta_tail)
        b.w_dt = _ptr(w_img) if fused else None
        b.bc32, b.bc32_tail = _ptr(job.get("bc32")), _ptr(job.get("bc32_tail"))
        b.idx, b.A, b.Dskip, b.dt_bias, b.ydir = _ptr(idxs[i]), _ptr(w["A"]), _ptr(w["Ds"]), _ptr(w["dt_bias"]), _ptr(job["ydir"])
        keep += [xdbl, xdbl_tail, delta, delta_tail, w_img, tail, job]
    if tc:
        gemm.run(dtproj, name="gemm_dtproj")
    elif jobs:
        fork.join()
    live = [i for i, n in enumerate(n_sels) if n > 0]
    if live:
        min_tiles = min(-(-(n_sels[i] + (0 if tails[i] is None else tails[i].shape[1])) // 16) for i in live)
        n_ctas = -(-Dk // 64) * Bp * 2 * len(live)
        n_ctas_shape = n_ctas if SCAN_BATCH_HINT is None else -(-Dk // 64) * SCAN_BATCH_HINT * 2 * len(live)
        args.nseg = SCAN_SEGMENTS if SCAN_SEGMENTS is not None else auto_segments(n_ctas_shape, min_tiles)
        if args.nseg <= 1:
            args.chain_chunks = SCAN_CHAIN if SCAN_CHAIN is not None else auto_chain(n_ctas, min_tiles)
        ws_bytes = lib.actk_masked_scan_workspace_bytes(ct.byref(args))
        if ws_bytes:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x0.device)
            args.workspace, args.workspace_bytes = ws.data_ptr(), ws_bytes
            keep.append(ws)
    if live:
        with torch.cuda.device(x0.device), _timed("masked_scan", x0.device):
            _lib.check(lib.actk_masked_scan_fwd(ct.byref(args), _stream(x0)), "actk_masked_scan_fwd")
    return outs


class SS2D_cond_v10(nn.Module):
    """mamba_layer.py:1902-1986, constructed exactly as TransformerSTmodel.py:3960-3974 does."""

    def __init__(self, d_model, d_cond, cond_size=0, d_state=16, d_conv=3, expand=2, dt_rank="auto",
                 dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4,
                 dropout=0.0, conv_bias=True, bias=False, device=None, dtype=None, size=8,
                 scan_type="scan", num_direction=8, **kwargs):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        unit_args = (d_model, d_cond, cond_size, d_state, d_conv, expand, dt_rank, dt_min, dt_max, dt_init,
                     dt_scale, dt_init_floor, dropout, conv_bias, bias, device, dtype, size, scan_type, num_direction)
        self.audio_unit = SS2D_Unit(*unit_args)
        self.exp_unit = SS2D_Unit(*unit_args)
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.d_cond = d_cond
        self.audio_proj = nn.Linear(d_cond, self.d_inner, bias=bias, **fk)
        self.exp_proj = nn.Linear(d_cond, self.d_inner, bias=bias, **fk)
        self.id_proj = nn.Linear(d_cond, self.d_inner, bias=bias, **fk)
        self.in_proj1 = nn.Linear(d_model, self.d_inner, bias=bias, **fk)
        self.in_proj2 = nn.Linear(d_model, self.d_inner, bias=bias, **fk)
        self.act1 = nn.SiLU()
        self.act2 = nn.SiLU()
        self.num_direction = num_direction
        self.out_norm = nn.LayerNorm(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else None   # never applied upstream either (:1952)
        self.scan_type = scan_type
        if scan_type != "sweep":
            # upstream's HSCANS_dynamic('scan') raises for every L > 1 (mamba_layer.py:150-151); only 'sweep'
            # (the identity order) is live (TransformerSTmodel.py:3969)
            raise NotImplementedError("only scan_type='sweep' is live in the reference")
        self.mask_cache = MaskIndexCache()

    def scan_core(self, xz1, xz2, tail1, tail2, m1: MaskIndex, m2: MaskIndex, ch_slice=None, weights=None,
                  layernorm=None, out_proj=False, push=None, out_buf=None, out_peers=None):
        """Both branches' gather -> bidirectional scan -> scatter, then direction/branch merge.
        ch_slice=None: + out_norm, returns (Bp, L, D) normalised.
        ch_slice=(lo, hi): returns the merged sums of channels [lo, hi) only, (Bp, L, hi-lo), NOT normalised —
        LayerNorm needs every channel and runs after the all-gather (sharded.py).
        out_proj=True: also apply out_proj and return (Bp, L, d_model) — fused into one tcgen05 kernel with the merge
        and LayerNorm where the shape is built (actk_merge_ln_outproj_supported), else merge kernel + cuBLAS.
        push=(peer_ptrs, my_part) with ch_slice: the merged slice is written straight into every rank's gather buffer
        over NVLink peer memory instead of a local tensor (sharded.py, gather="p2p"); returns None."""
        lib = _lib.load()
        Bp, L, D = xz1.shape
        res = _scan_branches([self.audio_unit, self.exp_unit], [xz1, xz2], [tail1, tail2], [m1.idx, m2.idx],
                             [m1.n_sel, m2.n_sel], Bp, L, idx64s=[m1.idx64, m2.idx64], ch_slice=ch_slice)
        Dk = res[0][1].shape[-1]
        out = None if push is not None else torch.empty((Bp, L, Dk), dtype=xz1.dtype, device=xz1.device)
        a = _lib.MergeLnArgs()
        for i, ((yd, xz_k), m) in enumerate(zip(res, (m1, m2))):
            a.xz[i], a.ydir[i], a.selected[i] = xz_k.data_ptr(), yd.data_ptr(), m.selected.data_ptr()
        a.out = _ptr(out)
        a.layernorm = (1 if ch_slice is None else 0) if layernorm is None else int(layernorm)
        if push is not None:
            ptrs, a.my_part = push
            a.n_peers = len(ptrs)
            for p_i, ptr in enumerate(ptrs):
                a.peer_out[p_i] = ptr
        if weights is not None:                      # v8 / v9: multiplicative blend with the downsampled masks
            wts = [w.to(xz1.dtype).contiguous() for w in weights]
            a.row_weight[0], a.row_weight[1] = wts[0].data_ptr(), wts[1].data_ptr()
        if a.layernorm:
            gamma, beta = self.out_norm.weight.to(xz1.dtype), self.out_norm.bias.to(xz1.dtype)
            a.gamma, a.beta = _ptr(gamma), _ptr(beta)
        a.eps, a.n_branches, a.Bp, a.L, a.D, a.dtype = self.out_norm.eps, 2, Bp, L, Dk, _DTYPES[xz1.dtype]
        w_out = self.out_proj.weight
        if (out_proj and a.layernorm and FUSE_LN_OUT_PROJ and self.out_proj.bias is None and w_out.dtype == xz1.dtype
                and lib.actk_merge_ln_outproj_supported(Dk, w_out.shape[0], a.dtype)):
            y = torch.empty((Bp, L, w_out.shape[0]), dtype=xz1.dtype, device=xz1.device)
            w_c = w_out if w_out.is_contiguous() else w_out.contiguous()
            with torch.cuda.device(xz1.device), _timed("merge_ln", xz1.device):
                _lib.check(lib.actk_merge_ln_outproj_fwd(ct.byref(a), _ptr(w_c), _ptr(y), w_out.shape[0], _stream(xz1)),
                           "actk_merge_ln_outproj_fwd")
            return y
        with torch.cuda.device(xz1.device), _timed("merge_ln", xz1.device):
            _lib.check(lib.actk_merge_layernorm_fwd(ct.byref(a), _stream(xz1)), "actk_merge_layernorm_fwd")
        return self._out_proj(out, out=out_buf, peers=out_peers) if out_proj else out

    def _out_proj(self, y, out=None, peers=None):
        """out_proj (mamba_layer.py:1985): tensor-core kernel for 16-bit activations, torch otherwise.
        out: optional (B', L, d_model) contiguous destination (a slot of a multi-GPU gather buffer).
        peers: device addresses of this call's (B'*L, d_model) slot in EVERY rank's gather buffer — the projection's
        epilogue then stores each tile to all of them over NVLink peer memory (GEMM + all-gather in one kernel) and
        nothing is returned."""
        w = self.out_proj.weight
        tc = TC_GEMM and self.out_proj.bias is None and gemm.usable(y, w) and y.shape[-1] % 8 == 0 and w.shape[0] % 8 == 0
        if peers is not None:
            if not tc or w.shape[0] % 64 != 0:
                raise NotImplementedError("the fused out_proj + all-gather needs the 16-bit tensor-core route and d_model % 64 == 0")
            a = y.reshape(-1, y.shape[-1])
            gemm.run([gemm.Problem(a if a.is_contiguous() else a.contiguous(), w if w.is_contiguous() else w.contiguous(),
                                   None, peers=peers)], name="gemm_outproj")
            return None
        if tc:
            return gemm.linear(y, w, name="gemm_outproj", out=out)
        res = self.out_proj(y)
        if out is not None:
            out.copy_(res)
            return out
        return res

    use_id = True   # SS2D_cond_v10_wo_id drops the identity token (and has no id_proj)

    def project_inputs(self, x, id_emb, conds, masks):
        """The dense front half of forward (mamba_layer.py:1958-1961, 1966, 1972, 1977): in_proj of both
        branches, id / condition projections, cached mask indices."""
        if not x.is_cuda:
            raise RuntimeError("actalker_b200 layers run on CUDA tensors only (no CPU path)")
        L = x.shape[1]
        # the index lists live on the activations' device wherever the mask tensors are (a CPU mask works upstream
        # too: indexing moves the index tensor); the downsample itself runs on the mask's own device and dtype
        m1 = self.mask_cache.get(masks[0], L, device=x.device)
        m2 = self.mask_cache.get(masks[1], L, device=x.device)
        Bp, n_c, dc = conds.shape
        D = self.d_inner
        if (self.use_id and n_c >= 2 and BATCH_IN_PROJ and self._tc_ok(x, id_emb, conds) and conds.is_contiguous()
                and x.is_contiguous()):
            # Tensor-core route, two launches on the current stream, no glue kernels:
            #   (1) audio_proj over ALL rows of `conds` + SiLU, stored one row further down in tail1, and exp_proj of the
            #       last token of every frame into slot 1 of tail2.  `conds` holds 32 audio tokens then the expression
            #       token per frame, tail1 the id token then the 32 audio tokens: the shift puts every audio token in
            #       place, and the rows it also fills with the (meaningless) audio projection of the expression token
            #       are exactly the id slots, which launch (2) overwrites;
            #   (2) in_proj1 | in_proj2 (x read once, two output planes) together with id_proj + SiLU into slot 0 of
            #       both tails.
            tail1 = torch.empty((Bp, n_c, D), dtype=x.dtype, device=x.device)
            tail2 = torch.empty((Bp, 2, D), dtype=x.dtype, device=x.device)
            gemm.run([gemm.Problem(conds.view(Bp * n_c, dc)[:-1], self.audio_proj.weight, tail1.view(Bp * n_c, D)[1:]),
                      gemm.Problem(conds[:, -1, :], self.exp_proj.weight, tail2[:, 1, :])], silu=True, name="gemm_cond")
            idm = id_emb.reshape(Bp, dc)
            ids = [gemm.Problem(idm, self.id_proj.weight, tail1[:, 0, :], silu=True),
                   gemm.Problem(idm, self.id_proj.weight, tail2[:, 0, :], silu=True)]
            xz1, xz2 = self._in_proj_both(x, extra=ids)
            return xz1, xz2, tail1, tail2, m1, m2
        fork = _Fork(x.device)
        with fork:                                   # 35 id / condition tokens per frame: side stream
            tail1, tail2 = self._tail_tokens(id_emb, conds)
        xz1, xz2 = self._in_proj_both(x)             # the latent tokens: current stream
        fork.join(tail1, tail2)
        return xz1, xz2, tail1, tail2, m1, m2

    def _tc_ok(self, *acts):
        """16-bit CUDA activations and weights of the same dtype, no biases, 16-byte rows: the tensor-core route."""
        ws = [self.in_proj1.weight, self.in_proj2.weight, self.audio_proj.weight, self.exp_proj.weight]
        if self.use_id:
            ws.append(self.id_proj.weight)
        lins = [self.in_proj1, self.in_proj2, self.audio_proj, self.exp_proj] + ([self.id_proj] if self.use_id else [])
        return (TC_GEMM and gemm.usable(*acts, *ws) and all(l.bias is None for l in lins)
                and self.d_model % 8 == 0 and self.d_cond % 8 == 0)

    def _tail_tokens(self, id_emb, conds):
        """tail1 = [SiLU(id_proj(id)), SiLU(audio_proj(audio tokens))], tail2 = [SiLU(id_proj(id)), SiLU(exp_proj(exp token))]
        (mamba_layer.py:1958-1960, 1966, 1977) — the route for fp32 activations and for the variant without an id token."""
        audio_cond, exp_cond = conds[:, :-1], conds[:, -1:]
        if not self.use_id and self._tc_ok(conds):           # v10_wo_id: no id slot to absorb a shifted store
            return (gemm.linear(audio_cond, self.audio_proj.weight, silu=True, name="gemm_cond"),
                    gemm.linear(exp_cond, self.exp_proj.weight, silu=True, name="gemm_cond"))
        tail1, tail2 = self.act1(self.audio_proj(audio_cond)), self.act2(self.exp_proj(exp_cond))
        if self.use_id:
            id_tok = self.act2(self.id_proj(id_emb))
            tail1, tail2 = torch.cat([id_tok, tail1], dim=1), torch.cat([id_tok, tail2], dim=1)
        return tail1.contiguous(), tail2.contiguous()

    def _in_proj_both(self, x, extra=()):
        """in_proj1(x), in_proj2(x) (mamba_layer.py:1960-1961) as ONE product: the two weights stacked to (2D, d_model),
        the two results written as separate contiguous planes — x is read once per call instead of once per branch.
        Tensor-core route (16-bit): one actk_gemm_tn_fwd launch; fp32: one batched torch GEMM over a stride-0 batch."""
        w1, w2 = self.in_proj1.weight, self.in_proj2.weight
        if (not BATCH_IN_PROJ or self.in_proj1.bias is not None or self.in_proj2.bias is not None or w1.dtype != x.dtype
                or w2.dtype != x.dtype):
            return self.in_proj1(x).contiguous(), self.in_proj2(x).contiguous()
        key = (w1.data_ptr(), w1._version, w2.data_ptr(), w2._version, w1.dtype, str(w1.device))
        if getattr(self, "_w_in_key", None) != key:
            with torch.no_grad():
                self._w_in = torch.stack([w1.t(), w2.t()], dim=0).contiguous()      # (2, d_model, D): torch.bmm form
                self._w_in_nk = torch.cat([w1, w2], dim=0).contiguous()             # (2D, d_model): nn.Linear form
            self._w_in_key = key
        Bp, L, dm = x.shape
        D = self.d_inner
        if self._tc_ok(x) and x.is_contiguous():
            xz = torch.empty((2, Bp * L, D), dtype=x.dtype, device=x.device)
            a = x.view(Bp * L, dm)
            if D % 32 == 0:
                gemm.run([gemm.Problem(a, self._w_in_nk, xz, planes=2), *extra], name="gemm_inproj")
            else:          # a column tile may not straddle the two planes: two problems of one launch
                gemm.run([gemm.Problem(a, w1, xz[0]), gemm.Problem(a, w2, xz[1]), *extra], name="gemm_inproj")
        else:
            assert not extra, "extra problems ride on the tensor-core launch only"
            x2 = x.reshape(1, Bp * L, dm).expand(2, Bp * L, dm)
            xz = torch.bmm(x2, self._w_in)                                           # (2, B'L, D)
        return xz[0].view(Bp, L, -1), xz[1].view(Bp, L, -1)

    def _check_forward_only(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError(f"actalker_b200.{type(self).__name__} is forward-only (the reference's "
                                      "inference path, pipeline ...two_ip.py:351); call it under torch.no_grad()")

    def forward(self, x, id_emb, conds, masks, out=None, out_peers=None):
        # x: (B', L, d_model); id_emb: (B', 1, d_cond); conds: (B', 33, d_cond) = 32 audio + 1 expression
        # tokens; masks: [audio (1,1,H,W), expression (1,1,H,W)]  (TransformerSTmodel.py:4121)
        # out (extension, optional): contiguous (B', L, d_model) destination, e.g. this rank's slot of a gather buffer
        self._check_forward_only(x)
        # out_peers (extension, optional): see _out_proj — the result goes to every rank's gather buffer instead
        return self.scan_core(*self.project_inputs(x, id_emb, conds, masks), out_proj=True, out_buf=out, out_peers=out_peers)


class SS2D_cond_v10_wo_id(SS2D_cond_v10):
    """mamba_layer.py:1988-2070: v10 without the identity token (no `id_proj` parameter; `id_emb` is ignored)."""
    use_id = False

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        del self.id_proj


def _full_index(m: MaskIndex, device) -> MaskIndex:
    """Index of ALL L tokens: v8 / v9 scan every token and blend afterwards."""
    iota = torch.arange(m.L, device=device)
    return MaskIndex(idx=iota.to(torch.int32), idx64=iota, selected=torch.ones(m.L, dtype=torch.uint8, device=device),
                     n_sel=m.L, L=m.L, weight=m.weight)


class SS2D_cond_v8(SS2D_cond_v10):
    """mamba_layer.py:1706-1800, the older multiplicative blend: both branches scan every latent token (plus the id
    and condition tail), each result is multiplied by its bicubically downsampled mask (:1777-1797), the two are
    added, then out_norm / out_proj.  Same parameters as v10; same kernels with per-row weights in the merge."""

    def _blend(self, x, id_emb, conds, masks, layernorm, out_proj=False):
        self._check_forward_only(x)
        xz1, xz2, tail1, tail2, m1, m2 = self.project_inputs(x, id_emb, conds, masks)
        f1, f2 = _full_index(m1, x.device), _full_index(m2, x.device)
        return self.scan_core(xz1, xz2, tail1, tail2, f1, f2, weights=[m1.weight, m2.weight], layernorm=layernorm,
                              out_proj=out_proj)

    def forward(self, x, id_emb, conds, masks):
        return self._blend(x, id_emb, conds, masks, layernorm=True, out_proj=True)


class SS2D_cond_v9(SS2D_cond_v8):
    """mamba_layer.py:1802-1899: v8 plus a third bidirectional unit (`fuse_unit`) over the blended sum before
    out_norm (:1893-1896)."""

    def __init__(self, d_model, d_cond, cond_size=0, d_state=16, d_conv=3, expand=2, dt_rank="auto", dt_min=0.001,
                 dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, dropout=0.0, conv_bias=True,
                 bias=False, device=None, dtype=None, size=8, scan_type="scan", num_direction=8, **kwargs):
        super().__init__(d_model, d_cond, cond_size, d_state, d_conv, expand, dt_rank, dt_min, dt_max, dt_init,
                         dt_scale, dt_init_floor, dropout, conv_bias, bias, device, dtype, size, scan_type,
                         num_direction, **kwargs)
        self.fuse_unit = SS2D_Unit(d_model, d_cond, cond_size, d_state, d_conv, expand, dt_rank, dt_min, dt_max,
                                   dt_init, dt_scale, dt_init_floor, dropout, conv_bias, bias, device, dtype, size,
                                   scan_type, num_direction)

    def forward(self, x, id_emb, conds, masks):
        y = self._blend(x, id_emb, conds, masks, layernorm=False)           # (B', L, D) blended sum
        y = self.fuse_unit(y.permute(0, 2, 1)).permute(0, 2, 1)
        return self.out_proj(self.out_norm(y))
