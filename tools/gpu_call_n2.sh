#!/bin/bash
# 2-GPU pass: the 2-GPU pytest, the peer-memory checks, bench.py under torchrun (weak value, e2e, strong block)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r9n2}
mkdir -p "$O"
timeout 600 python -m pytest tests/test_sharded.py -q --timeout 300 -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_batch_push.py 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 20 --warmup 5 > "$O/bench_n2.json" 2> "$O/bench_n2.err"
echo "bench_n2_rc=$?"
python - "$O/bench_n2.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],"link_frac",d["e2e"]["host_link_frac"])
print(json.dumps(d.get("strong"))[:1500])
PY
tail -3 "$O/bench_n2.err"
