#!/bin/bash
# GEMM table + bench split under library variants: VARIANTS="a b"
cd "$GRAFT_REPO_ROOT" || exit 1
show() { python -c "
import json,sys
for l in sys.stdin:
    try:
        d=json.loads(l)
        if '320' in d['product']: print('   ', d['product'][:60].ljust(62), d['ours_us'], d['cublas_us'])
    except Exception: pass
"; }
for v in default ${VARIANTS}; do
  echo "== $v"
  if [ "$v" = default ]; then unset ACTK_LIB_PATH; else export ACTK_LIB_PATH=$PWD/actalker_b200/lib/libactk_$v.so; fi
  timeout 200 python tools/bench_gemm_tn.py 2>/dev/null | show
  timeout 200 python bench.py --no-cpu-baseline --steps 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('    step', round(d['ms_per_step'],4), d['roofline']['ms_per_step_by_kernel'])"
done
