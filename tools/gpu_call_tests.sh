#!/bin/bash
# what the driver runs at round end: smoke(), the full GPU test suite, the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-final}
mkdir -p "$O"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > "$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?" | tee "$O/rc.txt"
tail -3 "$O/pytest_gpu.log"
timeout 600 python bench.py > "$O/bench_init.json" 2> "$O/bench_init.err"
echo "bench_rc=$?"
python - "$O/bench_init.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'frac', round(d['roofline']['frac'],4), 'scan', round(d['roofline']['kernel_ms'],4), 'parity', d['parity']['ok'], d['parity']['max_abs_err'], 'e2e', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value'],4), 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
