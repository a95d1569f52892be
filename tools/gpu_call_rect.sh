#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r8}
mkdir -p "$O"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > "$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?"; tail -3 "$O/pytest_gpu.log"
cd tools && python profile_kernels.py 320 50 rects 2>/dev/null; python profile_kernels.py 1280 50 rects 2>/dev/null | head -8; cd ..
timeout 300 python tools/bench_configs.py --only 3 > "$O/configs_3.jsonl" 2>/dev/null
python - "$O/configs_3.jsonl" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k!='per_layer'}, {k:(v['layer_ms'],v['scan_ms']) for k,v in d.get('per_layer',{}).items()})
PY
