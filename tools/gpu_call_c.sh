#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2c}
mkdir -p "$O"
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > "$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?" | tee -a "$O/rc.txt"
tail -4 "$O/pytest_gpu.log"
timeout 600 python bench.py > "$O/bench_init.json" 2> "$O/bench_init.err"
echo "bench_rc=$?" | tee -a "$O/rc.txt"
tail -c 1500 "$O/bench_init.json"
timeout 300 python tools/tune_variants.py run default poly0 poly1 poly3 > "$O/variants.txt" 2>&1
cat "$O/variants.txt"
timeout 120 tools/bin/microbench_pipes > "$O/microbench_pipes.txt" 2>&1
tail -16 "$O/microbench_pipes.txt"
