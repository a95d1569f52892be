#!/bin/bash
# full GPU test suite + a default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r7}
mkdir -p "$O"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > "$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?" | tee "$O/rc.txt"
tail -5 "$O/pytest_gpu.log"
timeout 600 python bench.py > "$O/bench_init.json" 2> "$O/bench_init.err"
echo "bench_rc=$?"
python - "$O/bench_init.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('parity'), d['e2e']['value'], d.get('ms_per_step_by_kernel'))
PY
