#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2n8}
mkdir -p "$O"
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus "$N" --steps 20 --warmup 5 > "$O/bench_n$N.json" 2> "$O/bench_n$N.err"
echo "bench_n${N}_rc=$?" | tee -a "$O/rc.txt"
python - "$O/bench_n$N.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],"link_frac",d["e2e"]["host_link_frac"])
    print(json.dumps(d.get("strong")))
except Exception as e: print("parse failed",e)
PY
done
# long clip (BASELINE configs[4]): F = 100 frames as ONE call over the 8 GPUs is the strong block's Bp100; the weak form:
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 10 --warmup 5 --frames 100 --no-strong > "$O/bench_n8_f100.json" 2> "$O/bench_n8_f100.err"
echo "bench_n8_f100_rc=$?" | tee -a "$O/rc.txt"
tail -c 600 "$O/bench_n8_f100.json"
