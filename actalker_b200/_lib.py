"""ctypes binding of the C-ABI library (include/actalker_b200.h).

There is no fallback: if libactalker_b200.so is missing or does not export every symbol the header
declares, importing a compute entry point raises.  The structures below mirror the header field by field.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ACTK_OK = 0
STATUS_NAMES = {0: "ACTK_OK", 1: "ACTK_ERR_BAD_SHAPE", 2: "ACTK_ERR_BAD_DTYPE", 3: "ACTK_ERR_BAD_ALIGN",
                4: "ACTK_ERR_BAD_ARG", 5: "ACTK_ERR_CUDA", 6: "ACTK_ERR_UNSUPPORTED"}
ACTK_F32, ACTK_F16, ACTK_BF16 = 0, 1, 2
ACTK_A_GENERAL, ACTK_A_POWER = 0, 1
ABI_VERSION = 12

EXPORTS = ["actk_abi_version", "actk_sm_arch", "actk_last_error", "actk_selective_scan_fwd",
           "actk_masked_scan_fwd", "actk_masked_scan_workspace_bytes", "actk_dt_proj_image_bytes",
           "actk_pack_dt_proj_weight", "actk_merge_ln_outproj_supported", "actk_merge_ln_outproj_fwd", "actk_merge_layernorm_fwd", "actk_gathered_layernorm_fwd", "actk_a_structure",
           "actk_scan_algorithmic_bytes", "actk_peer_buffer_alloc", "actk_peer_buffer_free", "actk_peer_buffer_export",
           "actk_peer_buffer_open", "actk_peer_buffer_close", "actk_gemm_tn_supported", "actk_gemm_tn_fwd", "actk_gather_rows"]

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class ScanArgs(C.Structure):
    _fields_ = [("u", _vp), ("delta", _vp), ("B", _vp), ("C", _vp), ("z", _vp),
                ("A", _vp), ("D", _vp), ("delta_bias", _vp),
                ("out", _vp), ("last_state", _vp),
                ("batch", _i), ("dim", _i), ("groups", _i), ("dstate", _i), ("seqlen", _i),
                ("u_sb", _ll), ("u_sd", _ll), ("delta_sb", _ll), ("delta_sd", _ll),
                ("z_sb", _ll), ("z_sd", _ll), ("out_sb", _ll), ("out_sd", _ll),
                ("B_sb", _ll), ("B_sg", _ll), ("B_sn", _ll), ("C_sb", _ll), ("C_sg", _ll), ("C_sn", _ll),
                ("dtype", _i), ("delta_softplus", _i), ("a_kind", _i)]


class BranchArgs(C.Structure):
    _fields_ = [("xz", _vp), ("tail", _vp), ("xdbl", _vp), ("xdbl_tail", _vp), ("delta", _vp),
                ("delta_tail", _vp), ("idx", _vp), ("A", _vp), ("Dskip", _vp), ("dt_bias", _vp), ("ydir", _vp),
                ("n_sel", _i), ("n_tail", _i), ("a_kind", _i), ("w_dt", _vp), ("bc32", _vp), ("bc32_tail", _vp)]


class MaskedScanArgs(C.Structure):
    _fields_ = [("br", BranchArgs * 2), ("n_branches", _i),
                ("Bp", _i), ("L", _i), ("D", _i), ("N", _i), ("xw", _i), ("dtype", _i),
                ("nseg", _i), ("workspace", _vp), ("workspace_bytes", _ll), ("chain_chunks", _i),
                ("dt_rank_pad", _i)]


class MergeLnArgs(C.Structure):
    _fields_ = [("xz", _vp * 2), ("ydir", _vp * 2), ("selected", _vp * 2),
                ("gamma", _vp), ("beta", _vp), ("out", _vp), ("eps", _f),
                ("n_branches", _i), ("Bp", _i), ("L", _i), ("D", _i), ("dtype", _i), ("layernorm", _i),
                ("row_weight", _vp * 2), ("peer_out", _vp * 8), ("n_peers", _i), ("my_part", _i)]


class GemmProblem(C.Structure):
    _fields_ = [("a", _vp), ("w", _vp), ("c", _vp), ("lda", _ll), ("ldw", _ll), ("ldc", _ll), ("plane_stride", _ll),
                ("M", _i), ("N", _i), ("K", _i), ("planes", _i), ("epilogue", _i), ("peer_c", _vp * 8), ("n_peers", _i),
                ("c_f32", _vp), ("ldc_f32", _ll), ("f32_cols", _i)]


GEMM_MAX_PROBLEMS = 4
GEMM_EPI_NONE, GEMM_EPI_SILU = 0, 1


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    # ACTK_LIB_PATH selects a tuning variant built by actalker_b200.build.build(out=..., defs=...)
    return os.environ.get("ACTK_LIB_PATH", _build.LIB_PATH)


def load():
    """Load (once) and type the shared library.  Raises LibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path) and "ACTK_LIB_PATH" not in os.environ:
        try:                      # fresh checkout: compile in-tree once (nvcc, sm_100a); still no non-CUDA path
            _build.build()
        except Exception as e:    # noqa: BLE001 - reported below
            raise LibraryMissing(f"{path} is missing and building it failed: {e}") from e
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not found: run `python -m actalker_b200.build` (nvcc, sm_100a). "
            "actalker_b200 has no CPU or PyTorch fallback for the scan path.")
    lib = C.CDLL(path)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise LibraryMissing(f"{path} does not export {missing}")
    lib.actk_abi_version.restype = _i
    lib.actk_sm_arch.restype = _i
    lib.actk_last_error.restype = C.c_char_p
    lib.actk_selective_scan_fwd.argtypes = [C.POINTER(ScanArgs), _vp]
    lib.actk_masked_scan_fwd.argtypes = [C.POINTER(MaskedScanArgs), _vp]
    lib.actk_masked_scan_workspace_bytes.argtypes = [C.POINTER(MaskedScanArgs)]
    lib.actk_masked_scan_workspace_bytes.restype = _ll
    lib.actk_dt_proj_image_bytes.argtypes = [_i, _i, _i]
    lib.actk_dt_proj_image_bytes.restype = _ll
    lib.actk_pack_dt_proj_weight.argtypes = [_vp, _i, _i, _i, _i, _vp, _vp]
    lib.actk_pack_dt_proj_weight.restype = _i
    lib.actk_merge_layernorm_fwd.argtypes = [C.POINTER(MergeLnArgs), _vp]
    lib.actk_merge_ln_outproj_supported.argtypes = [_i, _i, _i]
    lib.actk_merge_ln_outproj_supported.restype = _i
    lib.actk_merge_ln_outproj_fwd.argtypes = [C.POINTER(MergeLnArgs), _vp, _vp, _i, _vp]
    lib.actk_merge_ln_outproj_fwd.restype = _i
    lib.actk_a_structure.argtypes = [_vp, _i, _i, _f, _vp, _vp]
    lib.actk_gathered_layernorm_fwd.argtypes = [_vp, _i, _ll, _i, _vp, _vp, _f, _vp, _i, _vp]
    lib.actk_gathered_layernorm_fwd.restype = _i
    lib.actk_peer_buffer_alloc.argtypes = [_ll, C.POINTER(_vp)]
    lib.actk_peer_buffer_export.argtypes = [_vp, C.c_char_p]
    lib.actk_peer_buffer_open.argtypes = [C.c_char_p, C.POINTER(_vp)]
    lib.actk_peer_buffer_free.argtypes = [_vp]
    lib.actk_peer_buffer_close.argtypes = [_vp]
    for name in ("actk_peer_buffer_alloc", "actk_peer_buffer_export", "actk_peer_buffer_open", "actk_peer_buffer_free",
                 "actk_peer_buffer_close"):
        getattr(lib, name).restype = _i
    lib.actk_gemm_tn_supported.argtypes = [C.POINTER(GemmProblem), _i]
    lib.actk_gemm_tn_supported.restype = _i
    lib.actk_gemm_tn_fwd.argtypes = [C.POINTER(GemmProblem), _i, _i, _vp]
    lib.actk_gemm_tn_fwd.restype = _i
    lib.actk_gather_rows.argtypes = [_vp, _vp, _vp, _i, _i, _i, _ll, _vp]
    lib.actk_gather_rows.restype = _i
    lib.actk_scan_algorithmic_bytes.argtypes = [_i, _i, _i, _i, _i, _i]
    lib.actk_scan_algorithmic_bytes.restype = _ll
    for name in ("actk_selective_scan_fwd", "actk_masked_scan_fwd",
           "actk_pack_dt_proj_weight", "actk_merge_ln_outproj_supported", "actk_merge_ln_outproj_fwd", "actk_merge_layernorm_fwd", "actk_a_structure"):
        getattr(lib, name).restype = _i
    if lib.actk_abi_version() != ABI_VERSION:
        raise LibraryMissing(f"{path}: ABI version {lib.actk_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status: int, what: str):
    """Map a non-zero actk_status to the exception the reference stack would raise:
    shape/dtype/alignment violations -> RuntimeError (as mamba-ssm does), unsupported -> NotImplementedError."""
    if status == ACTK_OK:
        return
    msg = load().actk_last_error().decode("utf-8", "replace")
    text = f"{what}: {STATUS_NAMES.get(status, status)}: {msg}"
    if status == 6:
        raise NotImplementedError(text)
    raise RuntimeError(text)
