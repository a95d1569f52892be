"""In-tree nvcc build of the C-ABI library (sm_100a only).

    python -m actalker_b200.build            # builds actalker_b200/lib/libactalker_b200.so if stale
    python -m actalker_b200.build --force

The .so is git-ignored but travels to the GPU box with the repo snapshot; nothing is JIT-compiled at
import time and there is no fallback: if the library is missing, `actalker_b200._lib` raises.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libactalker_b200.so")
# (source, extra defines, object tag).  selective_scan.cu and merge_ln.cu are compiled once per I/O dtype
# (-DACTK_TU_DTYPE=n instantiates that dtype's kernels only) plus once for their C-ABI entry points, and the masked
# scan has one translation unit per dt-rank slab count, so that the long template instantiations build in parallel.
SOURCES = [("api.cu", (), ""), ("masked_scan.cu", (), ""),
           ("masked_scan_ks0.cu", (), ""), ("masked_scan_ks2.cu", (), ""), ("masked_scan_ks3.cu", (), ""),
           ("masked_scan_ks5.cu", (), ""), ("masked_scan_lean.cu", (), ""), ("ln_outproj.cu", (), ""), ("gemm_tn.cu", (), ""),
           ("selective_scan.cu", (), ""), ("merge_ln.cu", (), "")] + \
          [(src, (f"ACTK_TU_DTYPE={n}",), f"_t{n}") for src in ("selective_scan.cu", "merge_ln.cu") for n in (0, 1, 2)]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(PKG), "include", "actalker_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = None, defs=()) -> str:
    """Build the library.  `out`/`defs` produce a tuning variant (e.g. defs=["ACTK_POLY_STATES=3"]) beside the
    default one; select it at run time with ACTK_LIB_PATH."""
    variant = out is not None
    out = out or LIB_PATH
    if not variant and not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    log = []
    tag = os.path.splitext(os.path.basename(out))[0]
    procs = []
    for src, src_defs, otag in SOURCES:   # one nvcc per translation unit, all at once
        obj = os.path.join(LIB_DIR, (tag + "_" if variant else "") + src.replace(".cu", otag + ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in (*defs, *src_defs)], "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
        objs.append(obj)
    for src, pr in procs:
        so, se = pr.communicate()
        log.append(se)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{so}\n{se}")
    cmd = [_nvcc(), "-shared", "-o", out, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(LIB_DIR, "ptxas.log" if not variant else tag + "_ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return out


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
