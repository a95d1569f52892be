#!/bin/bash
# the round's single-GPU records: full GPU test suite, bench records (tools/profile_round.sh), ncu pass (tools/profile_round2.sh)
cd "$GRAFT_REPO_ROOT" || exit 1
O=${1:-r10}
mkdir -p "gpurun_out/$O"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > "gpurun_out/$O/pytest_gpu.log" 2>&1
echo "pytest_rc=$?"; tail -3 "gpurun_out/$O/pytest_gpu.log"
timeout 1500 bash tools/profile_round.sh "$O"
timeout 1500 bash tools/profile_round2.sh "${O}prof"
