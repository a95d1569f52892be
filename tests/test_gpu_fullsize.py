"""Parity at BASELINE.json's full sizes (config 2: B'=25, 72x72 tokens, d_model 320) where the CPU oracle would
take minutes: size-independent properties of the path, plus a full-size cross-check of the fused layer kernels
against the reference's own op graph (oracle layer code on the GPU) driven by the operator-contract kernel —
two different kernels, layouts and code paths that must agree."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import SS2D_cond_v10_ref

pytestmark = pytest.mark.gpu

B, SIDE, DM = 25, 72, 320
L = SIDE * SIDE


def make_layer(dtype, trained=True, seed=0):
    from actalker_b200 import SS2D_cond_v10
    torch.manual_seed(72589 + seed)
    layer = SS2D_cond_v10(d_model=DM, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=SIDE,
                          scan_type="sweep", num_direction=2).eval()
    if trained:
        with torch.no_grad():
            for u in (layer.audio_unit, layer.exp_unit):
                u.A_logs.add_(0.5 * torch.randn_like(u.A_logs))
                u.Ds.copy_(1.0 + 0.2 * torch.randn_like(u.Ds))
    if dtype != torch.float32:
        layer = layer.to(dtype)
        for n, p in layer.named_parameters():
            if any(s in n for s in ("A_logs", "Ds", "dt_projs_bias")):
                p.data = p.data.float()
    return layer.cuda()


def inputs(dtype, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, L, DM, device="cuda", generator=g).to(dtype)
    idm = torch.randn(B, 1, 1024, device="cuda", generator=g).to(dtype)
    cd = torch.randn(B, 33, 1024, device="cuda", generator=g).to(dtype)
    return x, idm, cd


def masks(kind, dtype):
    ones = torch.ones(1, 1, 576, 576, device="cuda", dtype=dtype)
    if kind == "ones":
        return [ones, ones.clone()]
    mouth = torch.zeros_like(ones); mouth[:, :, 330:480, 180:400] = 1      # like test_preprocess.py:259-267
    upper = torch.zeros_like(ones); upper[:, :, 60:330, 100:480] = 1
    return [mouth, upper]


@pytest.mark.parametrize("kind", ["ones", "rects"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_layer_equals_reference_graph_on_operator_kernel(dtype, kind):
    from actalker_b200 import selective_scan_fn
    ours = make_layer(dtype)
    ref = SS2D_cond_v10_ref(d_model=DM, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=SIDE,
                            scan_type="sweep", num_direction=2).eval()
    if dtype != torch.float32:
        ref = ref.to(dtype)
    ref = ref.cuda()
    for n, p in ref.named_parameters():
        p.data = dict(ours.named_parameters())[n].data.clone()
    x, idm, cd = inputs(dtype)
    m = masks(kind, dtype)
    with torch.no_grad():
        got = ours(x, idm, cd, m)
        want = ref(x.clone(), idm, cd, m, selective_scan=selective_scan_fn)
    n_sel = [ours.mask_cache.get(mm, L).n_sel for mm in m]
    assert n_sel == ([L, L] if kind == "ones" else n_sel) and all(0 < n <= L for n in n_sel)
    rtol, atol = (1e-3, 1e-4) if dtype == torch.float32 else (3e-2, 3e-2)
    err = (got.float() - want.float()).abs()
    assert torch.isfinite(got).all()
    assert (err <= atol + rtol * want.float().abs()).all(), err.max()


def test_power_path_equals_general_path_full_size():
    from actalker_b200 import _lib
    layer = make_layer(torch.bfloat16, trained=False)
    for u in (layer.audio_unit, layer.exp_unit):            # exact S4D structure in fp32
        u.A_logs.data = torch.log(torch.arange(1, 17, dtype=torch.float32, device="cuda")).repeat(2 * u.d_inner, 1)
    x, idm, cd = inputs(torch.bfloat16)
    m = masks("ones", torch.bfloat16)
    with torch.no_grad():
        assert layer.audio_unit.derived()["a_kind"] == _lib.ACTK_A_POWER
        y_pow = layer(x, idm, cd, m)
        for u in (layer.audio_unit, layer.exp_unit):
            u.derived()["a_kind"] = _lib.ACTK_A_GENERAL      # same weights through the general-A kernel
        y_gen = layer(x, idm, cd, m)
    assert torch.allclose(y_pow.float(), y_gen.float(), rtol=2e-2, atol=2e-2)
    assert (y_pow.float() - y_gen.float()).abs().mean() < 2e-3


def test_zero_masks_pass_in_proj_through_full_size():
    layer = make_layer(torch.bfloat16)
    x, idm, cd = inputs(torch.bfloat16)
    z = torch.zeros(1, 1, 576, 576, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        got = layer(x, idm, cd, [z, z.clone()])
        want = layer.out_proj(layer.out_norm(layer.in_proj2(x) + layer.in_proj1(x)))
    assert torch.allclose(got.float(), want.float(), rtol=2e-2, atol=2e-2)


def test_batch_permutation_is_exact_and_calls_are_deterministic():
    layer = make_layer(torch.bfloat16)
    x, idm, cd = inputs(torch.bfloat16)
    m = masks("rects", torch.bfloat16)
    perm = torch.randperm(B, device="cuda")
    with torch.no_grad():
        y1 = layer(x, idm, cd, m)
        y2 = layer(x, idm, cd, m)
        yp = layer(x[perm], idm[perm], cd[perm], m)
    assert torch.equal(y1, y2)
    assert torch.equal(yp, y1[perm])


def test_a_block_of_frames_reproduces_the_whole_call_when_the_launch_shape_is_pinned():
    """Frames are independent (mamba_layer.py:1955-1986 has no cross-frame operation), so a block of a call's frames must
    give the call's own rows.  Four frames alone would take the two-level scan (fp32 sums associate differently: one 16-bit
    rounding step of difference is allowed); with SCAN_BATCH_HINT = the whole call's frame count — what
    BatchShardedCall(exact=True) sets on every rank — the block is bit-identical."""
    from actalker_b200 import mamba_layer as ml
    layer = make_layer(torch.bfloat16)
    x, idm, cd = inputs(torch.bfloat16)
    m = masks("ones", torch.bfloat16)
    with torch.no_grad():
        whole = layer(x, idm, cd, m)
        free = layer(x[3:7], idm[3:7], cd[3:7], m)
        try:
            ml.SCAN_BATCH_HINT = B
            pinned = layer(x[3:7], idm[3:7], cd[3:7], m)
        finally:
            ml.SCAN_BATCH_HINT = None
    assert torch.equal(pinned, whole[3:7])
    assert torch.allclose(free.float(), whole[3:7].float(), rtol=3e-2, atol=3e-2)


def test_unit_direction_symmetry_full_size():
    """Swapping the two directions' weights and reversing the token order reverses the output
    (mamba_layer.py:1518-1547)."""
    from actalker_b200 import SS2D_Unit
    torch.manual_seed(3)
    a = SS2D_Unit(DM, 1024, 32, 16, size=SIDE, scan_type="sweep", num_direction=2).eval()
    with torch.no_grad():
        a.A_logs.add_(0.5 * torch.randn_like(a.A_logs))
    b = SS2D_Unit(DM, 1024, 32, 16, size=SIDE, scan_type="sweep", num_direction=2).eval()
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    D = a.d_inner
    for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias"):
        sd[k] = sd[k].flip(0)
    sd["A_logs"] = torch.cat([sd["A_logs"][D:], sd["A_logs"][:D]])
    sd["Ds"] = torch.cat([sd["Ds"][D:], sd["Ds"][:D]])
    b.load_state_dict(sd)
    a, b = a.cuda(), b.cuda()
    x = torch.randn(4, D, L + 37, device="cuda")               # odd length: ragged last tile in both directions
    with torch.no_grad():
        ya, yb = a(x), b(x.flip(-1))
    assert torch.allclose(yb, ya.flip(-1), rtol=1e-4, atol=1e-4)


def test_operator_linearity_in_u_full_size():
    from actalker_b200 import selective_scan_fn
    g = torch.Generator(device="cuda").manual_seed(5)
    Lp, Dm = 5217, 1280
    u1 = torch.randn(B, Dm, Lp, device="cuda", generator=g)
    u2 = torch.randn(B, Dm, Lp, device="cuda", generator=g)
    delta = torch.randn(B, Dm, Lp, device="cuda", generator=g)
    A = -torch.exp(torch.randn(Dm, 16, device="cuda", generator=g))
    Bm = torch.randn(B, 2, 16, Lp, device="cuda", generator=g)
    Cm = torch.randn(B, 2, 16, Lp, device="cuda", generator=g)
    Dv = torch.randn(Dm, device="cuda", generator=g)
    bias = torch.randn(Dm, device="cuda", generator=g) - 2
    f = lambda v: selective_scan_fn(v, delta, A, Bm, Cm, Dv, None, bias, True)
    lhs, rhs = f(u1 + 2 * u2), f(u1) + 2 * f(u2)
    scale = rhs.abs().max()
    assert ((lhs - rhs).abs() <= 1e-4 * scale + 1e-4 * rhs.abs()).all()


def test_bench_line_carries_the_contract_keys():
    """bench.py prints ONE JSON line with the driver's contract: metric / value / unit / timing fields, the roofline of the
    dominant kernel measured live, the end-to-end number through the public host-buffer API with its byte counts,
    clocks sampled during the timed region and the count of this repo's kernel launches."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "masked selective-scan Gtokens/s" and d["unit"] == "Gtokens/s" and d["n_gpus"] == 1
    assert d["steps"] == 3 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0.05 < rf["frac"] < 1.0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert abs(rf["achieved"] - rf["algorithmic_bytes"] / (rf["kernel_ms"] * 1e-3) / 1e9) < 1e-3 * rf["achieved"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.05
    # this repo's launches per step on the 16-bit route: cond projections, in_proj (+ id_proj), x_proj, dt_proj, the
    # masked scan, merge + LayerNorm, out_proj — all of them C-ABI launches of the in-tree library
    assert d["gpu_launches"] == 7 * d["steps"] and d["gpu_launches_per_step"] == 7
    per = rf["ms_per_step_by_kernel"]
    assert set(per) == {"gemm_cond", "gemm_inproj", "gemm_xproj", "gemm_dtproj", "masked_scan", "merge_ln", "gemm_outproj"}
    assert sum(per.values()) <= d["ms_per_step"] * 1.02           # the kernels of a step fit inside the step
    assert abs(d["value"] - 25 * 5184 / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6
    fl = rf["instruction_floor"]
    assert 0.5 < fl["frac_of_floor"] <= 1.05 and fl["ms"] < rf["kernel_ms"] * 1.05
    assert sorted(e["passes_ms"])[1] == round(e["ms_per_step"], 4) and e["reported_pass"] == "median"
    assert 0 < e["host_link_frac"] <= 1.05 and e["host_link"]["ms_per_step"] > 0


def test_bench_parity_field_compares_the_gpu_layer_with_the_cpu_oracle_on_the_bench_frames():
    """With the CPU baseline enabled the bench line carries `parity`: the GPU layer against the oracle on the very frames
    the cpu_baseline computed (full 72x72 frames of the bench workload), within the layer tolerance."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "3", "--warmup", "3", "--no-config0",
                        "--frames", "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    par, cb = d["parity"], d["cpu_baseline"]
    assert par["ok"] is True and par["finite"] is True and par["frames"] >= 1
    assert par["max_abs_err"] <= par["tol"]["atol"] + par["tol"]["rtol"] * par["max_abs_ref"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0


def test_reference_op_graph_on_the_upstream_derived_kernel_full_size(capsys):
    """The reference's own op graph (oracle layer code, op for op: scatter passes, flipped copies, einsums) on the GPU,
    driven by the mamba-ssm-derived CUDA kernel that ships in vLLM — the nearest runnable stand-in for
    `SS2D_cond_v10` + mamba-ssm on this box — against the drop-in layer at BASELINE config 2.  Checks parity and that
    the drop-in is faster; the measured ratio is printed (pytest -s) and recorded in DESIGN.md."""
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn as vllm_scan
    except Exception as e:   # noqa: BLE001
        pytest.skip(f"vllm's mamba kernel is not importable here: {type(e).__name__}")
    dtype = torch.bfloat16

    def upstream_scan(u, delta, A, Bm, Cm, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
        state = torch.zeros(u.shape[0], u.shape[1], A.shape[1], device=u.device, dtype=u.dtype)
        flag = torch.zeros(u.shape[0], dtype=torch.bool, device=u.device)
        return vllm_scan(u.contiguous(), state, delta.contiguous().clone(), A, Bm.contiguous(), Cm.contiguous(), D, z,
                         delta_bias, delta_softplus, has_initial_state=flag)

    ours = make_layer(dtype)
    ref = SS2D_cond_v10_ref(d_model=DM, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=SIDE,
                            scan_type="sweep", num_direction=2).eval().to(dtype).cuda()
    for n, p in ref.named_parameters():
        p.data = dict(ours.named_parameters())[n].data.clone()
    x, idm, cd = inputs(dtype)
    m = masks("ones", dtype)
    with torch.no_grad():
        try:
            want = ref(x.clone(), idm, cd, m, selective_scan=upstream_scan)
        except Exception as e:   # noqa: BLE001
            pytest.skip(f"vllm selective_scan_fwd not runnable here: {type(e).__name__}: {str(e)[:120]}")
        got = ours(x, idm, cd, m)
        err = (got.float() - want.float()).abs()
        assert (err <= 3e-2 + 3e-2 * want.float().abs()).all(), err.max()

        def t(fn, n=5):
            fn(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                fn()
            e.record(); torch.cuda.synchronize()
            return s.elapsed_time(e) / n
        t_ref = t(lambda: ref(x.clone(), idm, cd, m, selective_scan=upstream_scan))
        t_ours = t(lambda: ours(x, idm, cd, m))
    with capsys.disabled():
        print(f"\n[config 2, bf16] reference op graph + mamba-ssm-derived kernel: {t_ref:.2f} ms/layer; drop-in: {t_ours:.2f} ms "
              f"({t_ref / t_ours:.1f}x)")
    assert t_ours < t_ref
