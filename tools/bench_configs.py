#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that are not the bench.py headline (configs[1]):

  config 3   full UNet Mamba stack: 5 x (d_model 320, 72x72) + 5 x (640, 36x36) + 5 x (1280, 18x18) layer calls
             at CFG x2 (B'=50, as BASELINE words it) and CFG x4 (B'=100, what the reference pipeline runs,
             pipeline ...two_ip.py:712), all-ones masks and a mouth / upper-face rectangle split
  config 5   long clip: F = 25..100 frames at 72x72, d_model 320 — per-frame layout B'=F (what the reference does)
             and the flattened layout B'=1, L'=F*5184 (two-level chunk + carry scan)

One JSON line per measurement (device-resident inputs, CUDA events, 3 warm-up + 5 timed calls).
    python tools/bench_configs.py [--only 3|5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from actalker_b200 import SS2D_cond_v10, _lib
from actalker_b200 import mamba_layer as ml
from actalker_b200.mask import MaskIndex

PEAK = 6557.1


def make_layer(d_model, dtype=torch.bfloat16, seed=0):
    torch.manual_seed(72589 + seed)
    layer = SS2D_cond_v10(d_model=d_model, d_cond=1024, cond_size=32, dropout=0.1, d_state=16,
                          size=int(72 / (d_model / 320)), scan_type="sweep", num_direction=2).eval().to(dtype)
    for n, p in layer.named_parameters():
        if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
            p.data = p.data.float()
    return layer.cuda()


def timed(fn, warm=3, it=5):
    for _ in range(warm):
        fn()
    ml.TIMING = {}
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    ev, ml.TIMING = ml.TIMING.get("events", []), None
    scan = sum(a.elapsed_time(b) for n, a, b in ev if n == "masked_scan") / it
    return s.elapsed_time(e) / it, scan


def scan_q(Bp, lens, D, es=2):
    q = _lib.load().actk_scan_algorithmic_bytes
    return sum(q(Bp, l, 2 * D, 2, 16, es) for l in lens)


def config3():
    dtype = torch.bfloat16
    ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
    mouth = torch.zeros_like(ones); mouth[:, :, 330:480, 180:400] = 1
    upper = torch.zeros_like(ones); upper[:, :, 60:330, 100:480] = 1
    layers = {dm: make_layer(dm) for dm in (320, 640, 1280)}
    for cfg in (2, 4):
        Bp = 25 * cfg
        for mname, masks in (("all-ones", [ones, ones.clone()]), ("mouth/upper rectangles", [mouth, upper])):
            total_ms = total_scan = 0.0
            total_q = 0
            per = {}
            for dm, layer in layers.items():
                side = int(72 / (dm / 320)); L = side * side
                x = torch.randn(Bp, L, dm, device="cuda").to(dtype)
                idm = torch.randn(Bp, 1, 1024, device="cuda").to(dtype)
                cd = torch.randn(Bp, 33, 1024, device="cuda").to(dtype)
                with torch.no_grad():
                    ms, scan = timed(lambda: layer(x, idm, cd, masks))
                    n = [layer.mask_cache.get(m, L).n_sel for m in masks]
                q = scan_q(Bp, [n[0] + 33, n[1] + 2], 2 * dm)
                per[dm] = {"L": L, "n_sel": n, "layer_ms": round(ms, 3), "scan_ms": round(scan, 3),
                           "scan_GBps": round(q / scan / 1e6, 1)}
                total_ms += 5 * ms; total_scan += 5 * scan; total_q += 5 * q
                del x, idm, cd
            print(json.dumps({"config": 3, "cfg": cfg, "Bp": Bp, "masks": mname, "stack_layers": 15,
                              "stack_ms": round(total_ms, 3), "stack_scan_ms": round(total_scan, 3),
                              "scan_algorithmic_GB": round(total_q / 1e9, 3),
                              "scan_roofline_frac": round(total_q / total_scan / 1e6 / PEAK, 4), "per_layer": per}))


def config5():
    dtype = torch.bfloat16
    layer = make_layer(320)
    D = 640
    ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
    for F in (25, 50, 75, 100):
        L = 5184
        x = torch.randn(F, L, 320, device="cuda").to(dtype)
        idm = torch.randn(F, 1, 1024, device="cuda").to(dtype)
        cd = torch.randn(F, 33, 1024, device="cuda").to(dtype)
        with torch.no_grad():
            ms, scan = timed(lambda: layer(x, idm, cd, [ones, ones]))
        q = scan_q(F, [L + 33, L + 2], D)
        print(json.dumps({"config": 5, "layout": "per-frame B'=F", "frames": F, "tokens": F * L, "layer_ms": round(ms, 3),
                          "scan_ms": round(scan, 3), "scan_roofline_frac": round(q / scan / 1e6 / PEAK, 4),
                          "Gtokens_per_s": round(F * L / ms / 1e6, 4)}))
        del x, idm, cd
        # flattened: one sequence of F*5184 tokens through the same scan core (two-level scan cuts it into chunks)
        Lf = F * L
        xz1 = torch.randn(1, Lf, D, device="cuda").to(dtype)
        xz2 = torch.randn(1, Lf, D, device="cuda").to(dtype)
        t1 = torch.randn(1, 33, D, device="cuda").to(dtype)
        t2 = torch.randn(1, 2, D, device="cuda").to(dtype)
        iota = torch.arange(Lf, device="cuda")
        m = MaskIndex(idx=iota.int(), idx64=iota, selected=torch.ones(Lf, dtype=torch.uint8, device="cuda"), n_sel=Lf, L=Lf)
        with torch.no_grad():
            ms, scan = timed(lambda: layer.scan_core(xz1, xz2, t1, t2, m, m))
        q = scan_q(1, [Lf + 33, Lf + 2], D)
        nseg = ml.auto_segments(10 * 1 * 4, (Lf + 2 + 15) // 16)
        print(json.dumps({"config": 5, "layout": "flattened B'=1", "frames": F, "tokens": Lf, "segments": nseg,
                          "scan_core_ms": round(ms, 3), "scan_ms_3_launches": round(scan, 3),
                          "scan_roofline_frac": round(q / scan / 1e6 / PEAK, 4)}))
        del xz1, xz2


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=0)
    a = ap.parse_args()
    if a.only in (0, 3):
        config3()
    if a.only in (0, 5):
        config5()
