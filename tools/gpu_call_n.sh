#!/bin/bash
# multi-GPU pass: bench.py under torchrun at N ranks (weak value, e2e + host-link ceiling, strong block)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out/${2:-r2n$N}
mkdir -p "$O"
shift; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus "$N" --steps 20 --warmup 5 "$@" > "$O/bench_n$N.json" 2> "$O/bench_n$N.err"
echo "bench_n${N}_rc=$?" | tee -a "$O/rc.txt"
tail -c 2500 "$O/bench_n$N.json"; tail -5 "$O/bench_n$N.err"
