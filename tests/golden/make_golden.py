#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REAL reference layer.

Runs only in the build container (needs /root/reference, read-only).  The
reference module src/models/base/mamba_layer.py imports timm, diffusers,
pyzorder and mamba_ssm at module top (:7, :10, :20-23, :45); none is installed
here, so they are stubbed: the names the live classes never touch become
placeholders, and the two third-party functions on the hot path are bound to
the restatements under oracle/ (selective_scan_fn -> oracle.selective_scan_ref,
IPAdapterMaskProcessor.downsample -> oracle.downsample).  Everything else —
SS2D_cond_v10.forward (:1955-1986), SS2D_Unit.forward_core (:1505-1548),
HSCANS_dynamic (:142-184), the parameter initialisers (:1450-1502) — is the
reference's own code.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import downsample, selective_scan_ref  # noqa: E402

REF_FILE = "/root/reference/src/models/base/mamba_layer.py"
SEED = 72589  # the reference's own seed, config/inference.yaml:133


def load_reference_module():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    class _MaskProc:
        downsample = staticmethod(downsample)

    stub("timm"); stub("timm.models"); stub("timm.models.resnet", Bottleneck=object)
    stub("timm.models.layers", DropPath=nn.Identity, to_2tuple=lambda x: (x, x), trunc_normal_=lambda *a, **k: None)
    stub("diffusers"); stub("diffusers.image_processor", IPAdapterMaskProcessor=_MaskProc)
    stub("mamba_ssm"); stub("mamba_ssm.ops")
    stub("mamba_ssm.ops.selective_scan_interface", selective_scan_fn=selective_scan_ref,
         selective_scan_ref=selective_scan_ref)
    stub("pyzorder", ZOrderIndexer=object)
    spec = importlib.util.spec_from_file_location("ref_mamba_layer", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rect_mask(hw, r0, r1, c0, c1, soft=False):
    m = torch.zeros(1, 1, hw, hw)
    m[:, :, r0:r1, c0:c1] = 1.0
    if soft:  # soft edge like the LANCZOS-resized masks of test_preprocess.py:284-285
        k = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0]) / 16.0
        m = torch.nn.functional.conv2d(m, (k[:, None] * k[None, :])[None, None], padding=2)
    return m


def make_layer_case(ref, name, d_model, d_cond, side, batch, masks, dtype=torch.float32, trained_like=False,
                    seed_off=0, cls="SS2D_cond_v10"):
    torch.manual_seed(SEED + seed_off)
    L = side * side
    layer = getattr(ref, cls)(d_model=d_model, d_cond=d_cond, cond_size=32, dropout=0.1, d_state=16,
                              size=side, scan_type="sweep", num_direction=2).eval()
    if trained_like:
        with torch.no_grad():
            for unit in [m for n, m in layer.named_children() if n.endswith("_unit")]:
                unit.A_logs.add_(0.5 * torch.randn_like(unit.A_logs))
                unit.Ds.copy_(1.0 + 0.2 * torch.randn_like(unit.Ds))
    if dtype != torch.float32:
        layer = layer.to(dtype)
        for pname, p in layer.named_parameters():          # Inference.py:430-433
            if any(s in pname for s in ("A_logs", "Ds", "dt_projs_bias")):
                p.data = p.data.to(torch.float32)
    x = torch.randn(batch, L, d_model).to(dtype)
    id_emb = torch.randn(batch, 1, d_cond).to(dtype)
    conds = torch.randn(batch, 33, d_cond).to(dtype)
    masks = [m.to(dtype) for m in masks]
    with torch.no_grad():
        y = layer(x.clone(), id_emb, conds, masks)
        idx = [ref.IPAdapterMaskProcessor.downsample(m[:, 0], m.shape[0], L, 1).view(-1).int().nonzero().view(-1)
               for m in masks]
    out = {"meta": np.array([d_model, d_cond, side, batch], dtype=np.int64),
           "dtype": np.array(str(dtype).replace("torch.", "")), "cls": np.array(cls),
           "x": x.float().numpy(), "id_emb": id_emb.float().numpy(), "conds": conds.float().numpy(),
           "mask0": masks[0].float().numpy(), "mask1": masks[1].float().numpy(),
           "idx0": idx[0].numpy(), "idx1": idx[1].numpy(), "y": y.float().numpy()}
    for k, v in layer.state_dict().items():
        out["sd." + k] = v.float().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: y{tuple(y.shape)} n_sel={[int(i.numel()) for i in idx]} |y|max={y.float().abs().max():.4f}")


def make_unit_case(ref, name, d_model, L, batch, seed_off):
    torch.manual_seed(SEED + seed_off)
    unit = ref.SS2D_Unit(d_model, 64, 32, 16, size=8, scan_type="sweep", num_direction=2).eval()
    with torch.no_grad():
        unit.A_logs.add_(0.3 * torch.randn_like(unit.A_logs))
        unit.dt_projs_bias.add_(torch.randn_like(unit.dt_projs_bias))
        x = torch.randn(batch, unit.d_inner, L)
        y = unit(x)
    out = {"meta": np.array([d_model, L, batch], dtype=np.int64), "x": x.numpy(), "y": y.numpy()}
    for k, v in unit.state_dict().items():
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: y{tuple(y.shape)}")


def main():
    ref = load_reference_module()
    ones = torch.ones(1, 1, 64, 64)
    zeros = torch.zeros(1, 1, 64, 64)
    make_layer_case(ref, "layer_ones_f32", 32, 64, 8, 2, [ones, ones], seed_off=1)
    make_layer_case(ref, "layer_rect_f32", 32, 64, 8, 3, [rect_mask(64, 32, 56, 16, 48), rect_mask(64, 0, 32, 0, 64)],
                    trained_like=True, seed_off=2)
    make_layer_case(ref, "layer_zero_soft_f32", 48, 64, 8, 2, [zeros, rect_mask(64, 8, 56, 8, 56, soft=True)],
                    trained_like=True, seed_off=3)
    make_layer_case(ref, "layer_ones_bf16", 32, 64, 8, 2, [ones, ones], dtype=torch.bfloat16, seed_off=4)
    make_layer_case(ref, "layer_rect_f16", 32, 64, 8, 2, [rect_mask(64, 32, 56, 16, 48), ones],
                    dtype=torch.float16, trained_like=True, seed_off=5)
    make_unit_case(ref, "unit_f32", 24, 77, 2, seed_off=6)
    # the older / ablation variants that run on the same kernels
    soft = rect_mask(64, 8, 56, 8, 56, soft=True)
    hard = rect_mask(64, 32, 56, 16, 48)
    make_layer_case(ref, "v10woid_rect_f32", 32, 64, 8, 2, [hard, ones], trained_like=True, seed_off=7, cls="SS2D_cond_v10_wo_id")
    make_layer_case(ref, "v8_soft_f32", 32, 64, 8, 2, [soft, hard], trained_like=True, seed_off=8, cls="SS2D_cond_v8")
    make_layer_case(ref, "v8_soft_bf16", 32, 64, 8, 2, [soft, ones], dtype=torch.bfloat16, seed_off=9, cls="SS2D_cond_v8")
    make_layer_case(ref, "v9_soft_f32", 32, 64, 8, 2, [soft, hard], trained_like=True, seed_off=10, cls="SS2D_cond_v9")


if __name__ == "__main__":
    main()
