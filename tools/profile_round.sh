set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
python bench.py > gpurun_out/r2/bench_init.json 2> gpurun_out/r2/bench_init.err
python bench.py --params trained --no-cpu-baseline > gpurun_out/r2/bench_trained.json 2>/dev/null
python bench.py --params s4d --no-cpu-baseline > gpurun_out/r2/bench_s4d.json 2>/dev/null
python bench.py --cfg 4 --no-cpu-baseline > gpurun_out/r2/bench_cfg4.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_reference.json 2>/dev/null
ACTK_FUSE_DT=1 python bench.py --no-cpu-baseline > gpurun_out/r2/bench_fused_dt.json 2>/dev/null
ACTK_FUSE_LN_OUT=1 python bench.py --no-cpu-baseline > gpurun_out/r2/bench_fused_lnout.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:masked_scan -s 6 -c 1 -o gpurun_out/r2/masked_scan_general python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:masked_scan -s 6 -c 1 -o gpurun_out/r2/masked_scan_power python bench.py --steps 2 --warmup 3 --no-cpu-baseline --params s4d > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:merge_ln -s 3 -c 1 -o gpurun_out/r2/merge_ln python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ACTK_FUSE_DT=1 ncu --set full --clock-control none --import-source on -k regex:masked_scan -s 6 -c 1 -o gpurun_out/r2/masked_scan_fused_dt python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/bench_configs.py > gpurun_out/r2/configs_3_5.jsonl 2>/dev/null
python tools/bench_latency.py > gpurun_out/r2/latency_graph.jsonl 2>/dev/null
python tools/bench_operator.py > gpurun_out/r2/operator.txt 2>/dev/null
ls -la gpurun_out/r2
