#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2d}
mkdir -p "$O"
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q --timeout 90 > "$O/gemm_test.log" 2>&1
echo "gemm_test_rc=$?" | tee -a "$O/rc.txt"
tail -3 "$O/gemm_test.log"
for mh in 1 2; do
ACTK_GEMM_MH=$mh timeout 240 python tools/bench_gemm_tn.py > "$O/gemm_bench_mh$mh.jsonl" 2> "$O/gemm_bench_mh$mh.err"
echo "mh=$mh"
python - "$O/gemm_bench_mh$mh.jsonl" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try:
        d=json.loads(l); print(d["product"][:60].ljust(62), d["ours_us"], d["cublas_us"], d["frac_of_measured_hbm_peak"], d["max_abs_diff_vs_cublas"])
    except Exception: pass
PY
done
timeout 300 python bench.py --no-cpu-baseline > "$O/bench_tc.json" 2> "$O/bench_tc.err"
echo "bench_tc_rc=$?" | tee -a "$O/rc.txt"
python - "$O/bench_tc.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], json.dumps(d["roofline"]["ms_per_step_by_kernel"]))
PY
