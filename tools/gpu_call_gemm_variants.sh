#!/bin/bash
# the GEMM table (products matching FILTER) under each ";"-separated environment setting of CONFIGS
cd "$GRAFT_REPO_ROOT" || exit 1
show() { python -c "
import json,sys
for l in sys.stdin:
    try:
        d=json.loads(l)
        if '${FILTER:-320}' in d['product']: print('   ', d['product'][:60].ljust(62), d['ours_us'], d['cublas_us'], d['max_abs_diff_vs_cublas'])
    except Exception: pass
"; }
IFS=';' read -ra CFGS <<< "${CONFIGS:-}"
for c in "${CFGS[@]}"; do
  echo "== $c"
  env $c timeout 150 python tools/bench_gemm_tn.py 2>/dev/null | show
done
