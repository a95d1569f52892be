#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2b}
mkdir -p "$O"
timeout 240 python -m pytest tests/test_gpu_gemm.py -x -q --timeout 90 > "$O/gemm_test.log" 2>&1
echo "gemm_test_rc=$?" | tee -a "$O/rc.txt"
tail -3 "$O/gemm_test.log"
timeout 240 python tools/bench_gemm_tn.py > "$O/gemm_bench.jsonl" 2> "$O/gemm_bench.err"
echo "gemm_bench_rc=$?" | tee -a "$O/rc.txt"
timeout 300 python bench.py --no-cpu-baseline > "$O/bench_tc.json" 2> "$O/bench_tc.err"
echo "bench_tc_rc=$?" | tee -a "$O/rc.txt"
python - "$O/gemm_bench.jsonl" <<'PY'
import json,sys
for l in open(sys.argv[1] if len(sys.argv)>1 else "'"$O"'/gemm_bench.jsonl"):
    try:
        d=json.loads(l); print(d["product"][:60].ljust(62), d["ours_us"], d["cublas_us"], d["frac_of_measured_hbm_peak"], d["max_abs_diff_vs_cublas"])
    except Exception: pass
PY
