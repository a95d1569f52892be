#!/bin/bash
# quick check: the given tests, then the by-kernel split of the default bench
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest ${TESTS:-tests/test_gpu_parity.py} -x -q --timeout 300 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --steps 10 ${EXTRA} | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['ms_per_step_by_kernel'])"
