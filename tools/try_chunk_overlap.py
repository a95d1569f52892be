#!/usr/bin/env python
"""Does splitting one B' = 100 call (CFG x4, pipeline ...two_ip.py:712) into frame chunks on alternating streams let the
memory-bound kernels of one chunk (projections, merge + LayerNorm) run under the compute-bound scan of another?
    python tools/try_chunk_overlap.py [B'] [chunks] [streams]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench_configs import make_layer

Bp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dtype = torch.bfloat16
layer = make_layer(320)
L = 5184
x = torch.randn(Bp, L, 320, device="cuda").to(dtype)
idm = torch.randn(Bp, 1, 1024, device="cuda").to(dtype)
cd = torch.randn(Bp, 33, 1024, device="cuda").to(dtype)
ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
masks = [ones, ones.clone()]


def timed(fn, it=5):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


def chunked(nchunks, nstreams, prio=False):
    main = torch.cuda.current_stream()
    streams = [torch.cuda.Stream(priority=(-1 if (prio and i % 2) else 0)) for i in range(nstreams)]
    bounds = [Bp * i // nchunks for i in range(nchunks + 1)]

    def run():
        for s in streams:
            s.wait_stream(main)
        outs = []
        for c in range(nchunks):
            lo, hi = bounds[c], bounds[c + 1]
            with torch.cuda.stream(streams[c % nstreams]):
                outs.append(layer(x[lo:hi], idm[lo:hi], cd[lo:hi], masks))
        for s in streams:
            main.wait_stream(s)
        return outs
    return run


with torch.no_grad():
    ref = layer(x, idm, cd, masks)
    print(f"B'={Bp} one call                      {timed(lambda: layer(x, idm, cd, masks)):.3f} ms")
    for nchunks, nstreams in [(2, 1), (4, 1), (2, 2), (4, 2), (4, 4), (8, 2), (8, 4)]:
        if nchunks > Bp:
            continue
        run = chunked(nchunks, nstreams)
        got = torch.cat(run(), dim=0)
        ok = torch.equal(got, ref)
        print(f"B'={Bp} {nchunks} chunks on {nstreams} stream(s)        {timed(run):.3f} ms  identical={ok}")
