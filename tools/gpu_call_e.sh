#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2e}
mkdir -p "$O"
timeout 300 python tools/tune_variants.py run default nofastsp poll1000 > "$O/variants.txt" 2>&1
cat "$O/variants.txt"
timeout 300 python tools/tune_variants.py run default nofastsp --params s4d > "$O/variants_s4d.txt" 2>&1
cat "$O/variants_s4d.txt"
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > "$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?" | tee -a "$O/rc.txt"
tail -4 "$O/pytest_gpu.log"
