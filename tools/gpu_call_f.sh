#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2f}
mkdir -p "$O"
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > "$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?" | tee -a "$O/rc.txt"
tail -4 "$O/pytest_gpu.log"
for bn in 128 192 256; do
echo "bn_max=$bn"
ACTK_GEMM_BN_MAX=$bn timeout 240 python tools/bench_gemm_tn.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['product'][:60].ljust(62), d['ours_us'], d['cublas_us'], d['frac_of_measured_hbm_peak'])
    except Exception: pass
"
done
timeout 300 python tools/bench_configs.py --only 3 > "$O/configs_3.jsonl" 2>/dev/null
python - "$O/configs_3.jsonl" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k!='per_layer'}, {k:(v['layer_ms'],v['scan_ms']) for k,v in d.get('per_layer',{}).items()})
PY
