#!/bin/bash
# time tuning variants of the library on the config-2 bench: VARIANTS="a b c" [LEAN=0|1] [EXTRA="--params s4d"]
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r6}
mkdir -p "$O"
for lean in ${LEAN:-0}; do
  echo "ACTK_LEAN_SCAN=$lean"
  ACTK_LEAN_SCAN=$lean timeout 600 python tools/tune_variants.py run default ${VARIANTS} ${EXTRA} | tee -a "$O/variants.txt"
done
