#!/bin/bash
# A/B of two builds of the library: GEMM tests on the default build, then the by-kernel split of the default bench for
# the default build and for each library named in LIBS (paths relative to the repo root)
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest ${TESTS:-tests/test_gpu_gemm.py} -x -q --timeout 90 2>&1 | tail -4
split() { timeout 200 python bench.py --no-cpu-baseline --steps 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['ms_per_step_by_kernel'])"; }
for rep in 1 2; do
  echo "== default build"; split
  for l in ${LIBS}; do echo "== $l"; ACTK_LIB_PATH="$GRAFT_REPO_ROOT/$l" split; done
done
