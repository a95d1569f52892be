#!/usr/bin/env python
"""Benchmark of the masked selective-scan layer (ACTalker SS2D_cond_v10) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cfg 1|4] [--params init|trained|s4d]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward of the layer (both branches, both directions, mask gather/scatter, merge + LayerNorm,
in/out projections) over one batch of synthetic input.  Workload at N=1 = BASELINE.json configs[1]:
B'=25 frames x 72x72 latent tokens, d_model 320 (D=640, N=16, K=2), bf16, all-ones region masks (what the
shipped Inference.py:545-546 feeds), parameters from the reference initialisers (mamba_layer.py:1450-1502).
With N>1 every rank runs that workload on its own frames (the batch x CFG axis shards with no collective):
weak scaling.  One JSON line is printed by rank 0.

  value     latent tokens/s through the layer with inputs resident in HBM (device-timed, max over ranks)
  e2e       same through the public module call from pinned HOST buffers, H2D of x/id/conds and D2H of y timed
            (median of three passes); e2e.host_link: what the box's host link alone sustains for those copies
  roofline  the dominant kernel (actk_masked_scan_fwd) alone: algorithmic bytes Q / its mean
            launch duration measured with CUDA events inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restated selective_scan_ref path) on the host cores, a bounded sample of the same
                workload, plus BASELINE configs[0] (median of 3); `--impl reference` times that path alone
  parity    the GPU layer against the CPU oracle on the very frames the cpu_baseline computed
  strong    (N>1) ONE call of B'=100 (CFG x4) and of B'=25 split batch-first over the ranks, all-gather of the result
            inside the timed region, against the same call on one GPU: efficiency t1 / (N * tN)
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 72589  # reference seed, config/inference.yaml:133
METRIC = "masked selective-scan Gtokens/s"
LAYER_TOL = {"f32": (1e-3, 1e-4), "bf16": (3e-2, 3e-2), "f16": (5e-3, 5e-3)}   # tests/test_gpu_parity.py LAYER_TOL


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def floor_constants():
    """Issue-port cost of one warp-step of the scan's hot loop, from the tracked microbenchmark record."""
    with open(os.path.join(ROOT, "profiles", "floor_constants.json")) as f:
        return json.load(f)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.nv = self.h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        if self.nv is None:
            return
        try:
            nv, h = self.nv, self.h
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.reasons |= {n for bit, n in names.items() if r & bit}
                time.sleep(0.002)
        except Exception as e:
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cast_like_inference(layer, dtype, params):
    """The reference's own flow: the whole UNet is cast to 16 bit (Inference.py:200-202) and A_logs / Ds / dt_projs_bias
    are cast BACK to fp32 (:430-433), so they hold 16-bit-rounded values ("init" / "trained").  "s4d": they never leave
    fp32, so A keeps the exact S4D-real structure A[d][n] = -(n+1)."""
    if dtype == torch.float32:
        return layer
    keep = {n: p.data.clone() for n, p in layer.named_parameters()
            if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias"))}
    layer = layer.to(dtype)
    for name, p in layer.named_parameters():
        if name in keep:
            p.data = keep[name] if params == "s4d" else p.data.float()
    return layer


def make_layer(cls, d_model, dtype, params, device, seed):
    torch.manual_seed(seed)
    layer = cls(d_model=d_model, d_cond=1024, cond_size=32, dropout=0.1, d_state=16,
                size=int(72 / (d_model / 320)), scan_type="sweep", num_direction=2).eval()
    if params == "trained":   # trained-like variant: breaks the S4D-real structure of A
        with torch.no_grad():
            for unit in (layer.audio_unit, layer.exp_unit):
                unit.A_logs.add_(0.5 * torch.randn_like(unit.A_logs))
                unit.Ds.copy_(1.0 + 0.2 * torch.randn_like(unit.Ds))
    return cast_like_inference(layer, dtype, params).to(device)


def host_inputs(Bp, L, d_model, dtype, seed, pin):
    """x (Bp, L, d_model), id_emb (Bp, 1, 1024), conds (Bp, 33, 1024) as views of ONE host buffer (one staging copy per
    step on the e2e path)."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(Bp, L, d_model), (Bp, 1, 1024), (Bp, 33, 1024)]
    sizes = [s[0] * s[1] * s[2] for s in shapes]
    pad = [(-n) % 64 for n in sizes]                     # keep every view 128-byte aligned inside the packed buffer
    buf = torch.empty(sum(sizes) + sum(pad), dtype=dtype)
    if pin:
        buf = buf.pin_memory()
    views, off = [], 0
    for s, n, p in zip(shapes, sizes, pad):
        v = buf[off:off + n].view(s)
        v.copy_(torch.randn(s, generator=g).to(dtype))
        views.append(v)
        off += n + p
    return views


def scan_bytes(Bp, L, D, es):
    # Q of SURVEY.md §8(d): both branches, all-ones masks: L'_audio = L+1+32, L'_exp = L+1+1
    from actalker_b200 import _lib
    q = _lib.load().actk_scan_algorithmic_bytes
    return q(Bp, L + 33, 2 * D, 2, 16, es) + q(Bp, L + 2, 2 * D, 2, 16, es)


CPU_DTYPES = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


def cpu_baseline(frames, threads, side=72, d_model=320, dtype="bf16", params="init", keep=False):
    """The reference's CPU path (oracle port of the layer + selective_scan_ref) on a bounded sample of the bench
    workload: `frames` of its frames at side x side tokens, same d_model / dtype / branches / masks.  The token loop
    of selective_scan_ref is sequential Python over L' = side^2 + 33 steps, so a pass takes seconds per frame.
    keep=True also returns (layer, inputs, output) for the parity check of the GPU arm."""
    from oracle import SS2D_cond_v10_ref
    torch.set_num_threads(threads)
    dt_ = CPU_DTYPES[dtype]
    if dt_ == torch.float16:
        dt_ = torch.bfloat16     # CPU GEMMs in fp16 are not generally available; same width, same traffic (stated in `sample`)
    layer = make_layer(SS2D_cond_v10_ref, d_model, dt_, params, "cpu", SEED + 1)
    x, idm, cd = host_inputs(frames, side * side, d_model, dt_, SEED + 1, pin=False)
    ones = torch.ones(1, 1, 8 * side, 8 * side, dtype=dt_)
    with torch.no_grad():
        t0 = time.perf_counter()
        y = layer(x.clone(), idm, cd, [ones, ones])
        dt = time.perf_counter() - t0
    val = frames * side * side / dt / 1e9
    return (val, dt, layer, (x, idm, cd, ones), y) if keep else (val, dt)


def workload_name(Bp, frames, cfg, side, d_model, params):
    return (f"SS2D_cond_v10 layer forward, BASELINE configs[1]: B'={Bp} ({frames} frames x CFG {cfg}) x {side}x{side} "
            f"tokens, d_model {d_model}, d_state 16, 2 branches x 2 directions, all-ones masks, {params} parameters")


def config_dict(args, world, channel=False):
    """The `config` block, IDENTICAL for the GPU arm and the reference arm of one command line."""
    side = int(72 / (args.d_model / 320))
    Bp = args.frames * args.cfg
    par = (f"d_inner channel-sharded x{world}: replicated in_proj/x_proj, sliced scan, " +
           ("one NCCL all-gather of the merged slices" if args.gather == "nccl" else
            "merge kernel pushes the slices into every rank's gather buffer over NVLink peer memory + 4-byte barrier "
            "all-reduce") + " before out_norm/out_proj") if channel else \
        f"batch-sharded x{world} (each rank its own B'={Bp} frames), no collective"
    return {"workload": workload_name(Bp, args.frames, args.cfg, side, args.d_model, args.params),
            "tokens_per_step_per_gpu": Bp * side * side,
            "l2": "GPU arm: inputs rotate over 3 resident sets; the per-step working set (~1.5 GB) exceeds the 126 MB L2",
            "a_kind": "power" if args.params == "s4d" or args.dtype == "f32" and args.params == "init" else "general",
            "parallelism": par}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path (selective_scan_ref inside the restated
    SS2D_cond_v10; mamba-ssm's CUDA build cannot exist here) on the host cores, on THIS bench's workload.  Every step is
    the FULL workload (all B' frames: same config as the GPU arm) whenever steps + warm-up fit in ~900 s; otherwise a
    bounded sample of its frames, declared in `same_config` / `config.sample`."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    side = int(72 / (args.d_model / 320))
    Bp = args.frames * args.cfg
    # two calibration passes (they also warm torch): a pass costs a fixed Python-loop part plus a per-frame part
    # (B200 box, 16 cores: 1.5 s + 0.77 s per extra frame)
    _, t1 = cpu_baseline(1, threads, side, args.d_model, args.dtype, args.params)
    nsteps = max(1, args.steps + args.warmup)
    frames, per_frame = 1, 0.3 * t1
    if Bp > 1:
        k = min(Bp, 3)
        _, tk = cpu_baseline(k, threads, side, args.d_model, args.dtype, args.params)
        per_frame = max((tk - t1) / (k - 1), 0.05 * t1)
        t_full = t1 + (Bp - 1) * per_frame
        if nsteps * t_full <= args.reference_budget:
            frames = Bp
        else:
            frames = max(1, min(Bp, 1 + int((200.0 / nsteps - t1) / per_frame)))
    for _ in range(args.warmup):
        cpu_baseline(frames, threads, side, args.d_model, args.dtype, args.params)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_baseline(frames, threads, side, args.d_model, args.dtype, args.params)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = frames * side * side / (ms / 1e3) / 1e9
    full = frames == Bp
    sample = (f"{'all' if full else frames} of the workload's {Bp} frames per step ({side}x{side} tokens each, d_model "
              f"{args.d_model}, {'bf16 (stands in for f16 on the CPU)' if args.dtype == 'f16' else args.dtype} I/O with "
              f"fp32 scan arithmetic, 2 branches), oracle port of selective_scan_ref on {threads} threads")
    cfg = config_dict(args, world)
    if not full:
        cfg["sample"] = sample
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gtokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": cfg,
        "same_config": full, "frames_per_step": frames,
        # a sampled run only: what a full-workload step would cost by the two-point calibration (fixed Python-loop part +
        # per-frame part; conservative — the per-frame cost falls as frames are added)
        **({} if full else {"full_workload_estimate": {"ms_per_step": 1e3 * (t1 + (Bp - 1) * per_frame),
                                                       "value": Bp * side * side / (t1 + (Bp - 1) * per_frame) / 1e9}}),
        "cpu_baseline": {"value": val, "unit": "Gtokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gtokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def bind_to_gpu_cpus(local_rank):
    """N>1: run this rank on the CPU cores nearest its GPU (NVML's ideal affinity, cut to the cores the container
    allows) so that the pinned staging buffers of the e2e path are first touched on the GPU's own NUMA node."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        ideal = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus = ideal & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:   # noqa: BLE001  (no NVML, or a container without the call: keep the inherited affinity)
        pass


def host_link_ceiling(dev, h_in, h_out, steps, barrier):
    """What the host link alone sustains for one step's copies: the packed input buffer up and the result down on two
    streams, no compute, all ranks at once.  Returns ms per step (device-timed)."""
    d_in = torch.empty_like(h_in, device=dev)
    d_out = torch.empty_like(h_out, device=dev)
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(2):
        with torch.cuda.stream(up):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(down):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(dev)
    s, e_up, e_down = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    s.record(up)
    down.wait_event(s)
    for _ in range(steps):
        with torch.cuda.stream(up):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(down):
            h_out.copy_(d_out, non_blocking=True)
    e_up.record(up)
    e_down.record(down)
    torch.cuda.synchronize(dev)
    barrier()
    return max(s.elapsed_time(e_up), s.elapsed_time(e_down)) / steps


def parity_check(cpu_pack, dev, dtype_name, params):
    """The GPU layer on the exact frames, parameters and masks the CPU oracle just computed (bench's cpu_baseline)."""
    from actalker_b200 import SS2D_cond_v10
    ref, (x, idm, cd, ones), want = cpu_pack
    kw = dict(d_model=ref.d_model, d_cond=1024, cond_size=32, dropout=0.1, d_state=16,
              size=int(72 / (ref.d_model / 320)), scan_type="sweep", num_direction=2)
    ours = SS2D_cond_v10(**kw).eval()
    ours = cast_like_inference(ours, x.dtype, params)
    ours.load_state_dict(ref.state_dict(), strict=True)       # same keys, shapes and dtypes
    ours = ours.to(dev)
    with torch.no_grad():
        got = ours(x.to(dev), idm.to(dev), cd.to(dev), [ones.to(dev), ones.to(dev)]).float().cpu()
    want = want.float()
    rtol, atol = LAYER_TOL[dtype_name if x.dtype != torch.bfloat16 else "bf16"]
    err = (got - want).abs()
    return {"frames": int(x.shape[0]), "max_abs_err": err.max().item(), "max_abs_ref": want.abs().max().item(),
            "tol": {"rtol": rtol, "atol": atol}, "worst_excess_over_tol": (err - (atol + rtol * want.abs())).max().item(),
            "ok": bool((err <= atol + rtol * want.abs()).all()), "finite": bool(torch.isfinite(got).all()),
            "against": "oracle.SS2D_cond_v10_ref (selective_scan_ref port) on the same frames, parameters and masks"}


def strong_scaling(args, layer, dev, rank, world, dtype, barrier, steps):
    """ONE layer call split batch-first over the ranks: every rank holds the call's full inputs, computes its frames
    (ShardPlan("batch")) and the result is all-gathered to every rank INSIDE the timed region; t1 is the same call on one
    GPU (measured on every rank at once, max taken).  Two calls: the live caller's B' = 4 x 25 = 100 (pipeline
    ...two_ip.py:712) and B' = 25."""
    import torch.distributed as dist
    from actalker_b200.sharded import BatchShardedCall
    d_model, side = args.d_model, int(72 / (args.d_model / 320))
    L = side * side
    ones = torch.ones(1, 1, 576, 576, dtype=dtype, device=dev)
    masks = [ones, ones.clone()]
    out = {}
    for Bp in (100, 25):
        g = torch.Generator().manual_seed(SEED + 7)
        x = torch.randn(Bp, L, d_model, generator=g).to(dtype).to(dev)
        idm = torch.randn(Bp, 1, 1024, generator=g).to(dtype).to(dev)
        cd = torch.randn(Bp, 33, 1024, generator=g).to(dtype).to(dev)
        call = BatchShardedCall(layer, tiles=args.strong_tiles)
        push = BatchShardedCall(layer, gather="p2p") if d_model % 64 == 0 and dtype != torch.float32 else None
        res = {}
        with torch.no_grad():
            for name, fn in (("t1", lambda: layer(x, idm, cd, masks)), ("tN", lambda: call(x, idm, cd, masks)),
                             ("tP", (lambda: push(x, idm, cd, masks)) if push else None)):
                if fn is None:
                    continue
                for _ in range(3):
                    y = fn()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                s.record()
                for _ in range(steps):
                    y = fn()
                e.record()
                barrier()
                t = torch.tensor([s.elapsed_time(e) / steps], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res[name] = t.item()
                res[name + "_y"] = y
            # frames are independent, so the split call must reproduce the one-GPU call: bit for bit when both take the same
            # scan launch shape, within fp32 re-association when a rank's few frames take the two-level scan instead
            same = bool(torch.equal(res["t1_y"], res["tN_y"]))
            diff = (res["t1_y"].float() - res["tN_y"].float()).abs().max()
            same_push = bool(torch.equal(res["tN_y"], res["tP_y"])) if push else True   # same kernels, other transport
            # exact=True pins the unsplit call's scan launch shape on every rank: bit-identical whatever the split
            same_exact = bool(torch.equal(res["t1_y"], BatchShardedCall(layer, exact=True)(x, idm, cd, masks)))
        flag = torch.tensor([1 if same else 0, 1 if same_push else 0, 1 if same_exact else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        out[f"Bp{Bp}"] = {"ms_1gpu": res["t1"], "ms_Ngpu": res["tN"], "efficiency": res["t1"] / (world * res["tN"]),
                          "frames_per_rank": [hi - lo for lo, hi in call.plan(Bp).all_bounds()],
                          "gather_bytes_per_rank": Bp * L * d_model * x.element_size(),
                          "bit_identical_to_one_gpu": bool(flag[0].item()), "max_abs_diff_vs_one_gpu": diff.item(),
                          "bit_identical_to_one_gpu_with_exact_launch_shape": bool(flag[2].item()),
                          "phases_ms": call.phase_ms(),
                          "tiles_per_rank": call.tiles}
        if push:   # fused out_proj + all-gather over NVLink peer memory (TMA stores into every rank's buffer)
            out[f"Bp{Bp}"]["fused_push"] = {"ms_Ngpu": res["tP"], "efficiency": res["t1"] / (world * res["tP"]),
                                            "phases_ms": push.phase_ms(),
                                            "bit_identical_to_nccl_route": bool(flag[1].item())}
            push._peer.close()
        del x, idm, cd, res
    out["what"] = ("one call split batch-first (whole frames per rank), result all-gathered to every rank (NCCL over "
                   f"NVLink) inside the timed region; up to {args.strong_tiles} tiles per rank (>= 25 frames each) so the "
                   "gather of tile i runs under the compute of tile i+1; efficiency = t1 / (N * tN).  fused_push: the "
                   "out_proj kernel itself stores every output tile into all ranks' buffers over NVLink peer memory (TMA "
                   "stores, no collective launch; a 4-byte all-reduce orders readers behind writers)")
    return out


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from actalker_b200 import SS2D_cond_v10
    from actalker_b200 import mamba_layer as ml

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if args.chain is not None:
        ml.SCAN_CHAIN = args.chain
    dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[args.dtype]
    es = 4 if dtype == torch.float32 else 2
    d_model, side = args.d_model, int(72 / (args.d_model / 320))
    L, D = side * side, 2 * args.d_model
    Bp = args.frames * args.cfg
    layer = make_layer(SS2D_cond_v10, d_model, dtype, args.params, dev, SEED + 2)
    channel = world > 1 and args.shard == "channel"
    if channel:   # strong scaling of one layer call: every rank gets the same inputs, scans a d_inner slice
        from actalker_b200.sharded import ShardedSS2DCondV10
        inner, layer = layer, ShardedSS2DCondV10(layer, mode="channel", gather=args.gather)
    else:
        inner = layer
    ones = torch.ones(1, 1, 576, 576, dtype=dtype, device=dev)
    masks = [ones, ones.clone()]
    if world > 1 and not args.no_numa_bind:
        bind_to_gpu_cpus(local_rank)   # before the pinned buffers are first touched
    hx, hid, hcd = host_inputs(Bp, L, d_model, dtype, SEED + 2 + (0 if channel else rank), pin=True)
    hy = torch.empty(Bp, L, d_model, dtype=dtype).pin_memory()
    # rotate over several resident input sets so no step finds its inputs in L2 (126 MB)
    nrot = 3
    dsets = [[t.to(dev) + 0 * i for t in (hx, hid, hcd)] for i in range(nrot)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for i in range(max(3, args.warmup)):
            layer(*dsets[i % nrot], masks)
        a_kind = inner.audio_unit.derived()["a_kind"]
        # ---------------- device-resident timed region (value, roofline)
        sampler = ClockSampler(local_rank)
        sampler.start()
        # inside the timed region only the roofline's kernel is bracketed by CUDA events (on its launching stream): a pair
        # of event records costs the stream ~5 us, and bracketing all seven launches made the step 3 % slower than it is
        ml.TIMING = {"only": {"masked_scan"}}
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for i in range(args.steps):
            layer(*dsets[i % nrot], masks)
        end.record()
        barrier()
        sampler.stop_flag = True
        ms_total = start.elapsed_time(end)
        n_launches = ml.TIMING.get("launches", 0)
        events = ml.TIMING.get("events", [])
        kern = {}
        for name, s, e in events:
            kern.setdefault(name, []).append(s.elapsed_time(e))
        # the split of the step over ALL its launches: a second, separately instrumented pass of the same K steps
        ml.TIMING = {}
        for i in range(args.steps):
            layer(*dsets[i % nrot], masks)
        torch.cuda.synchronize(dev)
        split_events, ml.TIMING = ml.TIMING.get("events", []), None
        kern_all = {}
        for name, s, e in split_events:
            kern_all.setdefault(name, []).append(s.elapsed_time(e))
        # ---------------- end-to-end through the public host-buffer API: every step uploads its inputs from pinned
        # host memory and downloads its result; HostStreamedLayer overlaps step i+1's H2D and step i-1's D2H with
        # step i's compute (three streams, double-buffered staging) — all bytes still move inside the timed region
        from actalker_b200.host_api import HostStreamedLayer
        runner = HostStreamedLayer(layer)
        hys = [hy, torch.empty_like(hy).pin_memory()]
        for i in range(3):
            runner.submit(hx, hid, hcd, masks, hys[i % 2])
        runner.drain()
        # The host link of a shared box sees other tenants' traffic (observed: the same binary at 2.3 and 4.9 ms per
        # step minutes apart while the device-resident time stayed at 2.12 ms), so the K-step region is timed three
        # times; the MEDIAN pass is reported and all three are listed in e2e.passes_ms.
        e2e_passes = []
        for _ in range(3):
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            s2.record(runner.h2d)
            for i in range(args.steps):
                runner.submit(hx, hid, hcd, masks, hys[i % 2])
            e2.record(runner.d2h)
            runner.drain()
            barrier()
            e2e_passes.append((max(s2.elapsed_time(e2), 0.0), (time.perf_counter() - t0) * 1e3))
        ms_e2e_total, wall_e2e = sorted(e2e_passes)[1]
        packed = hx._base if hx._base is not None else hx
        link_ms = host_link_ceiling(dev, packed, hy, args.steps, barrier)
    sampler.join(timeout=1.0)

    ms_step, ms_e2e = ms_total / args.steps, ms_e2e_total / args.steps
    if world > 1:
        t = torch.tensor([ms_step, ms_e2e, link_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e, link_ms = t.tolist()
    strong = None
    if world > 1 and not channel and not args.no_strong:
        strong = strong_scaling(args, inner, dev, rank, world, dtype, barrier, max(3, min(args.steps, 10)))
    if rank != 0:
        return
    tokens = Bp * L * (1 if channel else world)
    peak, peak_src = peaks()
    q = scan_bytes(Bp, L, D // world if channel else D, es)   # per-rank launch
    scan_ms = statistics.mean(kern["masked_scan"])
    merge_ms = statistics.mean(kern_all["merge_ln"])
    per_step = {k: sum(v) / args.steps for k, v in kern_all.items()}        # ms per step, summed over the step's launches
    launches_per_step = n_launches / args.steps
    achieved = q / (scan_ms * 1e-3) / 1e9
    updates = Bp * (2 * L + 35) * 2 * D * 16
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and (Bp, d_model, args.dtype) == (25, 320, "bf16") and not channel:
        with open(tpath) as f:
            ent = json.load(f).get("masked_scan config2 bf16 " + {0: "general", 1: "power"}[a_kind])
        traffic = ent["bytes"] if ent else None
    fc = floor_constants()
    kind = {0: "general", 1: "power"}[a_kind]
    floor_cycles = fc["scan_hot_loop"][kind]["issue_cycles_per_warp_step"]
    floor_ms = (updates / 16 / 32) * floor_cycles / (148 * 4) / (fc["sm_clock_mhz"] * 1e6) * 1e3   # warp-steps over 592 schedulers
    cb = parity = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        # ~10-30 s of CPU work on a bounded sample of this workload: as many of its frames as ~15 s hold (a pass costs
        # a fixed Python-loop part t1 plus ~0.3 t1 per extra frame), at least one
        _, t1 = cpu_baseline(1, threads, side, d_model, args.dtype, args.params)
        nfr = max(1, min(Bp, 1 + int((15.0 - t1) / (0.3 * t1)))) if t1 < 15.0 else 1
        v, dt, ref_layer, ref_in, ref_out = cpu_baseline(nfr, threads, side, d_model, args.dtype, args.params, keep=True)
        cb = {"value": v, "unit": "Gtokens/s", "cores": threads, "kind": "port", "seconds": dt,
              "sample": f"{nfr} of the workload's {Bp} frames ({side}x{side} tokens each, d_model {d_model}, "
                        f"{'bf16 (stands in for f16 on the CPU)' if args.dtype == 'f16' else args.dtype} "
                        "I/O with fp32 scan arithmetic, 2 branches), one pass after a 1-frame warm-up pass"}
        parity = parity_check((ref_layer, ref_in, ref_out), dev, args.dtype, args.params)
        # configs[0] (B'=14 x 32x32, d_model 320, fp32) — 1 warm-up + median of 3
        if not args.no_config0:
            c0 = [cpu_baseline(14, threads, 32, 320, "f32", "init")[1] for _ in range(4)][1:]
            med = statistics.median(c0)
            cb["configs0"] = {"workload": "BASELINE configs[0]: B'=14 x 32x32 tokens, d_model 320, fp32, 2 branches, all-ones masks",
                              "seconds_median_of_3": med, "seconds": c0, "value": 14 * 1024 / med / 1e9, "unit": "Gtokens/s",
                              "scan_positions_per_s": 14 * (1057 + 1026) / med}
    out = {
        "metric": METRIC, "value": tokens / (ms_step * 1e-3) / 1e9, "unit": "Gtokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if channel else "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config_dict(args, world, channel),
        "roofline": {"bound": "hbm", "kernel": "masked_scan_kernel (actk_masked_scan_fwd)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes": q, "kernel_ms": scan_ms, "state_updates_per_s": updates / (scan_ms * 1e-3),
                     "merge_ln_ms": merge_ms, "kernel_share_of_step": scan_ms / ms_step,
                     "ms_per_step_by_kernel": {k: round(v, 5) for k, v in sorted(per_step.items())},
                     # The scan is bound by the SM's instruction issue, not by HBM, at d_state 16: every
                     # channel-step needs 16 exponentials and >= 64 fp32 multiply-adds.  profiles/floor_constants.json
                     # holds the issue cost of the hot loop's SASS instruction mix from the per-opcode rates measured
                     # on B200 (profiles/r01_microbench.txt, r02_microbench.txt): packed FFMA2 / FMUL2 hold the issue
                     # port 2.3 cycles, everything else 1.13, MUFU occupies the XU pipe 8.06.
                     "instruction_floor": {
                         "smsp_cycles_per_warp_step": floor_cycles,
                         "what": fc["scan_hot_loop"][kind]["what"],
                         "ms": floor_ms,
                         "frac_of_floor": floor_ms / scan_ms,
                         # the roofline fraction this kernel would show if it ran exactly at that floor: what
                         # fp32 per-(channel, state) exponentials allow on 148 SMs, whatever the memory system does
                         "frac_at_floor": (q / peak / 1e6) / floor_ms,
                         "source": fc["source"]},
                     # the same step's instruction mix run in isolation (no tiles, no HBM traffic) at saturating occupancy,
                     # next to the rate the kernel sustains over the whole launch (tile plumbing, chained chunks included)
                     "step_mix": {
                         "saturated_smsp_cycles_per_warp_step": fc["scan_hot_loop"][kind]["step_mix_saturated_cycles_per_warp_step"],
                         "kernel_smsp_cycles_per_warp_step": scan_ms * 1e-3 * fc["sm_clock_mhz"] * 1e6 * 148 * 4 / (updates / 16 / 32),
                         "what": fc["scan_hot_loop"][kind]["step_mix_what"]}},
        "e2e": {"value": tokens / (ms_e2e * 1e-3) / 1e9, "unit": "Gtokens/s",
                "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in (hx, hid, hcd)),
                "d2h_bytes_per_step": hy.numel() * hy.element_size(), "ms_per_step": ms_e2e,
                "api": "actalker_b200.host_api.HostStreamedLayer.submit (H2D | compute | D2H streams, depth 2; the three "
                       "inputs travel as one packed staging copy)",
                "host_wall_ms_per_step": wall_e2e / args.steps,
                "passes_ms": [round(p[0] / args.steps, 4) for p in e2e_passes], "reported_pass": "median",
                # the copies alone, no compute, all ranks at once: the ceiling the host link sets for this step
                "host_link": {"ms_per_step": link_ms, "ceiling_value": tokens / (link_ms * 1e-3) / 1e9,
                              "GBps_per_gpu_each_way": hy.numel() * hy.element_size() / (link_ms * 1e-3) / 1e9},
                "host_link_frac": link_ms / ms_e2e},
        "gpu_launches": int(round(launches_per_step * args.steps)),
        "gpu_launches_per_step": launches_per_step,
        "clocks": sampler.summary(),
    }
    if cb:
        out["cpu_baseline"] = cb
    if parity:
        out["parity"] = parity
    if strong:
        out["strong"] = strong
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--cfg", type=int, default=1)
    ap.add_argument("--d-model", dest="d_model", type=int, default=320)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16", "f32"])
    ap.add_argument("--params", default="init", choices=["init", "trained", "s4d"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config0", action="store_true", help="skip the BASELINE configs[0] CPU timing (4 passes of ~8 s)")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong-scaling block")
    ap.add_argument("--strong-tiles", type=int, default=2, help="N>1: tiles per rank whose all-gather overlaps the next tile")
    ap.add_argument("--reference-budget", type=float, default=900.0,
                    help="--impl reference: run the FULL workload per step when steps + warm-up fit in this many seconds")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: keep the inherited CPU affinity")
    ap.add_argument("--chain", type=int, default=None, help="force the number of chained chunks (tuning)")
    ap.add_argument("--shard", default="batch", choices=["batch", "channel"],
                    help="N>1: batch = weak scaling, no collective (default); channel = strong scaling of one call "
                         "with the NCCL all-gather before out_norm")
    ap.add_argument("--gather", default="nccl", choices=["nccl", "p2p"],
                    help="--shard channel: NCCL all-gather (default) or the merge kernel's fused push over NVLink peer memory")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
