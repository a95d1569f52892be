#!/bin/bash
# One GPU-box pass of round 2: the new tensor-core GEMM first (short timeouts: a hang must not eat the call), then the
# whole GPU suite, then bench lines with this repo's projections and with torch's for comparison.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2a}
mkdir -p "$O"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$O/smi.txt" 2>&1
timeout 240 python -m pytest tests/test_gpu_gemm.py -x -q --timeout 90 > "$O/gemm_test.log" 2>&1
echo "gemm_test_rc=$?" | tee -a "$O/rc.txt"
tail -5 "$O/gemm_test.log"
timeout 240 python tools/bench_gemm_tn.py > "$O/gemm_bench.jsonl" 2> "$O/gemm_bench.err"
echo "gemm_bench_rc=$?" | tee -a "$O/rc.txt"
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > "$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?" | tee -a "$O/rc.txt"
tail -8 "$O/pytest_gpu.log"
timeout 300 python bench.py --no-cpu-baseline > "$O/bench_tc.json" 2> "$O/bench_tc.err"
echo "bench_tc_rc=$?" | tee -a "$O/rc.txt"
ACTK_TC_GEMM=0 timeout 300 python bench.py --no-cpu-baseline > "$O/bench_cublas.json" 2> "$O/bench_cublas.err"
echo "bench_cublas_rc=$?" | tee -a "$O/rc.txt"
cat "$O/bench_tc.json" | head -c 600
