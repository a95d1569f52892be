"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/actalker_b200.h
declares, mirrors the header's structs, and rejects bad arguments before touching a device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from actalker_b200 import _lib


def header_text():
    with open(os.path.join(ROOT, "include", "actalker_b200.h")) as f:
        return f.read()


def test_library_builds_and_exports_every_declared_symbol():
    lib = _lib.load()
    text = re.sub(r"/\*.*?\*/", "", header_text(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(actk_[a-z0-9_]+)\s*\(", text)))
    assert declared, "no declarations parsed"
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.actk_abi_version() == _lib.ABI_VERSION == int(re.search(r"ACTK_ABI_VERSION (\d+)", text).group(1))
    assert lib.actk_sm_arch() == 100


def test_no_torch_types_in_the_boundary():
    text = header_text()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text and "#include <torch" not in text


def test_algorithmic_bytes_matches_survey_figures():
    lib = _lib.load()
    q = lambda b, l: lib.actk_scan_algorithmic_bytes(b, l, 1280, 2, 16, 2)
    assert abs((q(25, 5217) + q(25, 5186)) / 1e9 - 2.031) < 1e-3
    q1 = lambda l: lib.actk_scan_algorithmic_bytes(14, l, 1280, 2, 16, 4)
    assert abs((q1(1057) + q1(1026)) / 1e9 - 0.456) < 1e-3                # config 1 (fp32)


def _status(name):
    return {v: k for k, v in _lib.STATUS_NAMES.items()}[name]


def test_argument_validation_happens_before_any_launch():
    lib = _lib.load()
    assert lib.actk_selective_scan_fwd(None, None) == _status("ACTK_ERR_BAD_ARG")
    assert b"NULL" in lib.actk_last_error()
    a = _lib.ScanArgs()
    a.dtype = 7
    assert lib.actk_selective_scan_fwd(C.byref(a), None) == _status("ACTK_ERR_BAD_DTYPE")
    a.dtype = _lib.ACTK_BF16
    assert lib.actk_selective_scan_fwd(C.byref(a), None) == _status("ACTK_ERR_BAD_SHAPE")
    a.batch, a.dim, a.groups, a.dstate, a.seqlen = 1, 6, 4, 16, 8
    assert lib.actk_selective_scan_fwd(C.byref(a), None) == _status("ACTK_ERR_BAD_SHAPE")   # 6 % 4
    a.groups, a.dstate = 2, 128
    assert lib.actk_selective_scan_fwd(C.byref(a), None) == _status("ACTK_ERR_UNSUPPORTED")
    a.dstate = 16
    assert lib.actk_selective_scan_fwd(C.byref(a), None) == _status("ACTK_ERR_BAD_ARG")     # null tensors

    m = _lib.MaskedScanArgs()
    m.dtype, m.n_branches, m.N, m.Bp, m.L, m.D, m.xw = _lib.ACTK_BF16, 2, 16, 1, 64, 100, 104
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_SHAPE")       # 200-byte rows
    assert b"multiple of 8" in lib.actk_last_error()
    m.D, m.xw = 128, 100
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_ALIGN")       # 200-byte pitch
    m.xw, m.N = 104, 8
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_UNSUPPORTED")
    m.N = 16
    m.br[0].n_sel = 65
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_SHAPE")       # n_sel > L
    m.br[0].n_sel = 4
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_ARG")         # null tensors

    g = _lib.MergeLnArgs()
    g.dtype, g.n_branches, g.Bp, g.L, g.D = _lib.ACTK_F16, 2, 1, 4, 12
    assert lib.actk_merge_layernorm_fwd(C.byref(g), None) == _status("ACTK_ERR_BAD_SHAPE")   # D % 8
    assert lib.actk_a_structure(None, 4, 16, 1e-6, None, None) == _status("ACTK_ERR_BAD_ARG")


def test_check_maps_status_to_reference_exceptions():
    lib = _lib.load()
    a = _lib.ScanArgs()
    a.dtype, a.batch, a.dim, a.groups, a.dstate, a.seqlen = 0, 1, 4, 1, 200, 4
    with pytest.raises(NotImplementedError):
        _lib.check(lib.actk_selective_scan_fwd(C.byref(a), None), "scan")
    a.dstate = 16
    with pytest.raises(RuntimeError):
        _lib.check(lib.actk_selective_scan_fwd(C.byref(a), None), "scan")


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "actalker_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_cpu_tensors_are_refused_loudly():
    import torch
    from actalker_b200 import SS2D_cond_v10, selective_scan_fn
    u = torch.randn(1, 4, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        selective_scan_fn(u, u, -torch.ones(4, 16), torch.randn(1, 1, 16, 8), torch.randn(1, 1, 16, 8))
    layer = SS2D_cond_v10(d_model=32, d_cond=64, cond_size=32, dropout=0.1, d_state=16, size=8,
                          scan_type="sweep", num_direction=2).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.randn(1, 64, 32), torch.randn(1, 1, 64), torch.randn(1, 33, 64),
              [torch.ones(1, 1, 64, 64), torch.ones(1, 1, 64, 64)])


def test_workspace_and_gathered_layernorm_validation():
    lib = _lib.load()
    m = _lib.MaskedScanArgs()
    m.Bp, m.D, m.nseg = 1, 640, 1
    assert lib.actk_masked_scan_workspace_bytes(C.byref(m)) == 0
    m.nseg = 23
    assert lib.actk_masked_scan_workspace_bytes(C.byref(m)) == 1 * 4 * 23 * 640 * 33 * 4
    m.dtype, m.n_branches, m.N, m.L, m.xw = _lib.ACTK_BF16, 1, 16, 64, 104
    m.nseg = 5000
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_ARG")
    assert lib.actk_gathered_layernorm_fwd(None, 2, 10, 64, None, None, 1e-5, None, _lib.ACTK_BF16, None) \
        == _status("ACTK_ERR_BAD_ARG")
    assert lib.actk_gathered_layernorm_fwd(1 << 20, 2, 10, 60, 1 << 20, 1 << 20, 1e-5, 1 << 20, _lib.ACTK_BF16, None) \
        == _status("ACTK_ERR_BAD_SHAPE")           # Ds % 8
    assert lib.actk_gathered_layernorm_fwd(1 << 20, 2, 10, 64, 1 << 20, 1 << 20, 1e-5, 1 << 20, 9, None) \
        == _status("ACTK_ERR_BAD_DTYPE")


def test_fused_dt_proj_argument_validation():
    """dt_rank_pad (fused dt_proj, header (2)): only f16/bf16, only the built slab counts, and xw must hold the columns."""
    lib = _lib.load()
    m = _lib.MaskedScanArgs()
    m.dtype, m.n_branches, m.N, m.Bp, m.L, m.D, m.xw = _lib.ACTK_F32, 1, 16, 1, 64, 128, 128
    m.dt_rank_pad = 32
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_DTYPE")
    m.dtype, m.dt_rank_pad = _lib.ACTK_BF16, 64
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_UNSUPPORTED")
    m.dt_rank_pad, m.xw = 48, 128
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_SHAPE")   # needs 64 + 96 columns
    m.dt_rank_pad, m.xw = 32, 128
    m.br[0].n_sel = 64                       # live branch without pointers
    assert lib.actk_masked_scan_fwd(C.byref(m), None) == _status("ACTK_ERR_BAD_ARG")
    assert b"NULL" in lib.actk_last_error()


def test_derived_weight_layout_for_the_fused_dt_proj():
    """x_proj rows are regrouped [B_0|C_0|B_1|C_1 | dt_0 | dt_1] with each direction's rank zero-padded to a 16-multiple
    the in-kernel tcgen05.mma handles (20 -> 32, 40 -> 48, 80 -> 80)."""
    import torch
    from actalker_b200 import SS2D_Unit
    from actalker_b200.mamba_layer import _rank_pad
    assert [_rank_pad(r) for r in (4, 20, 32, 40, 80, 81)] == [(32, True), (32, True), (32, True), (48, True),
                                                              (80, True), (88, False)]
    u = SS2D_Unit(d_model=320, d_cond=64, cond_size=32, d_state=16, size=8, scan_type="sweep", num_direction=2)
    dv = u.derived()
    R, N, D, Rp = 20, 16, 640, 32
    assert dv["xw"] == 4 * N + 2 * Rp and dv["rank_pad"] == Rp and dv["fusable"]
    w = dv["w_xproj"]
    for k in range(2):
        assert torch.equal(w[k * 2 * N:(k + 1) * 2 * N], u.x_proj_weight[k, R:])
        assert torch.equal(w[4 * N + k * Rp:4 * N + k * Rp + R], u.x_proj_weight[k, :R])
        assert not w[4 * N + k * Rp + R:4 * N + (k + 1) * Rp].any()
        assert torch.equal(dv["w_dt"][k * Rp:k * Rp + R, k * D:(k + 1) * D], u.dt_projs_weight[k].t())


def test_gemm_argument_validation_and_struct_layout():
    """actk_gemm_tn_fwd (header (3c)) rejects bad launches before touching a device; the ctypes struct mirrors the header."""
    lib = _lib.load()
    text = header_text()
    body = re.search(r"typedef struct \{([^}]*)\} actk_gemm_problem;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = [n.strip(" *") for decl in body.split(";") if decl.strip() for n in decl.split(",")]
    declared = [re.sub(r"\[.*\]", "", n.split()[-1].strip("*")) for n in declared]
    assert declared == [f[0] for f in _lib.GemmProblem._fields_], declared
    assert int(re.search(r"ACTK_GEMM_MAX_PEERS (\d+)", text).group(1)) == 8 == len(_lib.GemmProblem().peer_c)
    assert int(re.search(r"ACTK_GEMM_MAX_PROBLEMS (\d+)", text).group(1)) == _lib.GEMM_MAX_PROBLEMS
    arr = (_lib.GemmProblem * 1)()
    assert lib.actk_gemm_tn_fwd(None, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_ARG")
    assert lib.actk_gemm_tn_fwd(arr, 0, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_ARG")
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_F32, None) == _status("ACTK_ERR_BAD_DTYPE")
    p = arr[0]
    p.a, p.w, p.c = 1 << 20, 1 << 21, 1 << 22
    p.M, p.N, p.K, p.planes, p.lda, p.ldw, p.ldc = 128, 64, 36, 1, 36, 36, 64
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_ALIGN")   # 72-byte rows
    assert not lib.actk_gemm_tn_supported(arr, _lib.ACTK_BF16)
    p.K = p.lda = p.ldw = 40
    assert lib.actk_gemm_tn_supported(arr, _lib.ACTK_BF16) and not lib.actk_gemm_tn_supported(arr, _lib.ACTK_F32)
    p.planes, p.N = 2, 72                       # 36 columns per plane: a tile would straddle the planes
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_SHAPE")
    p.planes, p.N, p.epilogue = 1, 64, 7
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_ARG")
    p.epilogue, p.n_peers = 0, 9                # fused all-gather: at most 8 ranks, non-NULL slots, N % 64 == 0
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_SHAPE")
    p.n_peers = 2
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_ARG") and b"peer" in lib.actk_last_error()
    p.peer_c[0], p.peer_c[1], p.N = 1 << 23, 1 << 24, 96
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_SHAPE")


def test_size_helpers_return_64_bit_values():
    """Workspace / image sizes are `long long` in the header; a c_int restype would truncate anything >= 2 GiB."""
    lib = _lib.load()
    assert lib.actk_masked_scan_workspace_bytes.restype is C.c_longlong
    assert lib.actk_dt_proj_image_bytes.restype is C.c_longlong
    m = _lib.MaskedScanArgs()
    m.Bp, m.D, m.nseg = 512, 2560, 400          # 512*4*400*2560*33*4 bytes = 276 GB: far above 2^31
    assert lib.actk_masked_scan_workspace_bytes(C.byref(m)) == 512 * 4 * 400 * 2560 * 33 * 4


def test_mask_index_lands_on_the_layers_device_and_weights_are_cached():
    """A mask may live on another device than the activations (the reference's indexing moves the index tensor); the
    cache key carries the target device.  Derived per-dtype weights are built once per parameter version."""
    import torch
    from actalker_b200 import SS2D_Unit
    from actalker_b200.mask import MaskIndexCache
    cache = MaskIndexCache()
    m = torch.zeros(1, 1, 64, 64)
    m[:, :, 16:48, 8:56] = 1
    e = cache.get(m, 64, device="cpu")
    assert e.idx.device.type == "cpu" and e.n_sel == 24 and e.selected.sum().item() == 24 and cache.misses == 1
    assert cache.get(m, 64, device=torch.device("cpu")) is e and cache.misses == 1
    with pytest.raises(RuntimeError, match="must be"):
        cache.get(torch.ones(64, 64), 64)
    u = SS2D_Unit(d_model=64, d_cond=64, cond_size=32, d_state=16, size=8, scan_type="sweep", num_direction=2)
    w = u.weights_for(torch.bfloat16)
    assert w is u.weights_for(torch.bfloat16) and w["w_xproj"].dtype == torch.bfloat16
    assert w["w_dt_nk"].shape == (2 * 128, 2 * u.derived()["rank_pad"]) and torch.equal(w["w_dt_nk"], w["w_dt"].t())
    half = u.weights_for(torch.bfloat16, 64, 128)           # channels [64, 128) of both directions
    assert half["A"].shape == (128, 16) and torch.equal(half["A"][:64], u.derived()["A"][64:128])
    assert torch.equal(half["w_dt_nk"][64:], w["w_dt_nk"][128 + 64:256])
    with torch.no_grad():
        u.Ds.add_(1.0)                                       # a parameter changed in place: everything is rebuilt
    assert u.weights_for(torch.bfloat16) is not w


def test_ctypes_structures_have_the_headers_layout(tmp_path):
    """Every argument structure of include/actalker_b200.h compiled by a plain C compiler (the header is C: no CUDA, no
    torch types) has the size and the field offsets of its ctypes mirror in actalker_b200/_lib.py."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"actk_scan_args": _lib.ScanArgs, "actk_branch_args": _lib.BranchArgs,
               "actk_masked_scan_args": _lib.MaskedScanArgs, "actk_merge_ln_args": _lib.MergeLnArgs,
               "actk_gemm_problem": _lib.GemmProblem}
    header = os.path.join(ROOT, "include", "actalker_b200.h")
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{header}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([cc, "-std=c99", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_gather_rows_and_fp32_side_output_validation():
    """actk_gather_rows and the fp32 side output of actk_gemm_tn_fwd reject bad arguments before any launch."""
    lib = _lib.load()
    assert lib.actk_gather_rows(None, None, None, 1, 1, 1, 16, None) == _status("ACTK_ERR_BAD_ARG")
    assert lib.actk_gather_rows(1 << 20, 1 << 21, 1 << 22, 1, 8, 4, 24, None) == _status("ACTK_ERR_BAD_ALIGN")
    assert lib.actk_gather_rows(1 << 20, 1 << 21, 1 << 22, 0, 8, 4, 16, None) == _status("ACTK_ERR_BAD_SHAPE")
    assert lib.actk_gather_rows(1 << 20, 1 << 21, 1 << 22, 2, 8, 0, 16, None) == _lib.ACTK_OK      # nothing selected: no launch
    arr = (_lib.GemmProblem * 1)()
    p = arr[0]
    p.a, p.w, p.c = 1 << 20, 1 << 21, 1 << 22
    p.M, p.N, p.K, p.planes, p.lda, p.ldw, p.ldc = 128, 128, 64, 1, 64, 64, 128
    p.f32_cols = 64                                   # no destination
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_ARG")
    p.c_f32, p.ldc_f32, p.f32_cols = 1 << 23, 64, 48  # 32 or 64 columns only
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) == _status("ACTK_ERR_BAD_SHAPE")
    p.f32_cols, p.ldc_f32 = 64, 32                    # pitch smaller than the columns
    assert lib.actk_gemm_tn_fwd(arr, 1, _lib.ACTK_BF16, None) in (_status("ACTK_ERR_BAD_SHAPE"), _status("ACTK_ERR_BAD_ALIGN"))
