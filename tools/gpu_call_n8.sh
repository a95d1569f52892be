#!/bin/bash
# 8-GPU pass: bench.py under torchrun (weak value, e2e + host-link ceiling, strong block), then the long clip of BASELINE
# configs[4] (F = 100 frames per rank, weak)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r12n8}
mkdir -p "$O"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --steps 20 --warmup 5 > "$O/bench_n8.json" 2> "$O/bench_n8.err"
echo "bench_n8_rc=$?"
python - "$O/bench_n8.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],"link_frac",d["e2e"]["host_link_frac"])
for k in ("Bp100","Bp25"):
    s=d["strong"][k]; print(k, s["ms_1gpu"], s["ms_Ngpu"], round(s["efficiency"],3), s["frames_per_rank"], s["bit_identical_to_one_gpu"], s["bit_identical_to_one_gpu_with_exact_launch_shape"], s["phases_ms"], "push", s["fused_push"]["ms_Ngpu"], round(s["fused_push"]["efficiency"],3), s["fused_push"]["bit_identical_to_nccl_route"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 10 --warmup 5 --frames 100 --no-strong > "$O/bench_n8_f100.json" 2> "$O/bench_n8_f100.err"
echo "bench_n8_f100_rc=$?"
python - "$O/bench_n8_f100.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("F=100 value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"])
PY
