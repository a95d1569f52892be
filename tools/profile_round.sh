#!/bin/bash
# The round's bench records on one B200 (no profiler): default line, parameter variants, CFG x4, the cuBLAS-route comparison,
# the opt-in fusions, a short reference arm, configs 3 / 5 and the operator seam.  tools/profile_round2.sh is the ncu pass.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2rec}
mkdir -p "$O"
python bench.py > "$O/bench_init.json" 2> "$O/bench_init.err"; echo "init_rc=$?"
python bench.py --params trained --no-cpu-baseline > "$O/bench_trained.json" 2>/dev/null; echo "trained_rc=$?"
python bench.py --params s4d --no-cpu-baseline > "$O/bench_s4d.json" 2>/dev/null; echo "s4d_rc=$?"
python bench.py --cfg 4 --no-cpu-baseline > "$O/bench_cfg4.json" 2>/dev/null; echo "cfg4_rc=$?"
ACTK_TC_GEMM=0 python bench.py --no-cpu-baseline > "$O/bench_cublas_route.json" 2>/dev/null; echo "cublas_rc=$?"
ACTK_FUSE_DT=1 python bench.py --no-cpu-baseline > "$O/bench_fused_dt.json" 2>/dev/null; echo "fused_dt_rc=$?"
ACTK_FUSE_LN_OUT=1 python bench.py --no-cpu-baseline > "$O/bench_fused_lnout.json" 2>/dev/null; echo "fused_ln_rc=$?"
python bench.py --dtype f16 --no-cpu-baseline > "$O/bench_f16.json" 2>/dev/null; echo "f16_rc=$?"
ACTK_LEAN_SCAN=1 python bench.py --no-cpu-baseline > "$O/bench_lean_scan.json" 2>/dev/null; echo "lean_rc=$?"
if [ -n "$REFERENCE_ARM" ]; then   # ~21 s of CPU per step on the 16-core box: only when asked for
python bench.py --impl reference --steps 3 --warmup 1 > "$O/bench_reference.json" 2>/dev/null; echo "reference_rc=$?"
fi
python tools/bench_gemm_tn.py > "$O/gemm_tn_vs_cublas.jsonl" 2>/dev/null; echo "gemm_rc=$?"
python tools/profile_gaps.py > "$O/timeline.txt" 2>/dev/null
python tools/bench_configs.py > "$O/configs_3_5.jsonl" 2>/dev/null; echo "configs_rc=$?"
python tools/bench_operator.py > "$O/operator.txt" 2>/dev/null; echo "operator_rc=$?"
python tools/bench_latency.py > "$O/latency_graph.jsonl" 2>/dev/null
python - "$O" <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(os.path.join(sys.argv[1],"bench_*.json"))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(os.path.basename(f).ljust(28), "ms/step", round(d["ms_per_step"],3), "value", round(d["value"],5), "scan", round(r.get("kernel_ms",0),3), "frac", round(r.get("frac",0),4), "e2e", round(d["e2e"]["ms_per_step"],3) if "ms_per_step" in d.get("e2e",{}) else "", d.get("same_config",""))
    except Exception as e: print(f, "ERR", e)
PY
