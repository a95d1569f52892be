#!/bin/bash
# lean scan kernel: parity, tuning variants, ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r4}
mkdir -p "$O"
timeout 600 python -m pytest tests/test_gpu_lean.py -x -q --timeout 300 > "$O/pytest_lean.log" 2>&1
echo "pytest_lean_rc=$?"; tail -3 "$O/pytest_lean.log"
timeout 400 python tools/tune_variants.py run default ${VARIANTS} > "$O/variants.txt" 2>&1
cat "$O/variants.txt"
ACTK_LEAN_SCAN=0 timeout 200 python tools/tune_variants.py run default | sed 's/default/general/'
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:masked_scan_lean -s 4 -c 1 -o "$O/masked_scan_lean" $B > /dev/null 2>&1
echo "ncu_rc=$?"
