#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
show() { python -c "
import json,sys
for l in sys.stdin:
    try:
        d=json.loads(l)
        if '320' in d['product']: print(d['product'][:70].ljust(72), d['ours_us'], d['cublas_us'], d['frac_of_measured_hbm_peak'], d['max_abs_diff_vs_cublas'])
    except Exception: pass
"; }
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lean.py -x -q --timeout 120 2>&1 | tail -2
timeout 200 python tools/bench_gemm_tn.py 2>/dev/null | show
timeout 200 python bench.py --no-cpu-baseline --steps 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['ms_per_step_by_kernel'])"
