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
for i in 1 2; do
echo default; timeout 200 python tools/bench_gemm_tn.py 2>/dev/null | show
echo epi3; ACTK_GEMM_EPI_BUFS=3 timeout 200 python tools/bench_gemm_tn.py 2>/dev/null | show
echo epi2; ACTK_GEMM_EPI_BUFS=2 timeout 200 python tools/bench_gemm_tn.py 2>/dev/null | show
done
