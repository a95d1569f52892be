#!/bin/bash
# GEMM tests, then the GEMM table under the environment settings given in CONFIGS (";"-separated "VAR=val VAR=val" groups)
cd "$GRAFT_REPO_ROOT" || exit 1
show() { python -c "
import json,sys
for l in sys.stdin:
    try:
        d=json.loads(l)
        if '${FILTER:-320}' in d['product']: print('   ', d['product'][:60].ljust(62), d['ours_us'], d['cublas_us'], d['max_abs_diff_vs_cublas'])
    except Exception: pass
"; }
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lean.py tests/test_gpu_unet_widths.py -x -q --timeout 120 2>&1 | tail -3
IFS=';' read -ra CFGS <<< "${CONFIGS:-}"
for c in "default" "${CFGS[@]}"; do
  echo "== $c"
  if [ "$c" = default ]; then timeout 200 python tools/bench_gemm_tn.py 2>/dev/null | show
  else env $c timeout 200 python tools/bench_gemm_tn.py 2>/dev/null | show; fi
done
