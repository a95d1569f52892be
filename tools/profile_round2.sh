#!/bin/bash
# Round-2 profile pass on one B200: ncu launch list of the bench command, `ncu --set full` of the scan, of the three large
# projection launches (in_proj, dt_proj, out_proj) and of the scan / merge launches at the UNet's other widths.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r2prof}
mkdir -p "$O"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > "$O/plain.log" 2>&1 || { echo "plain bench failed"; tail -5 "$O/plain.log"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/launches.csv" $B > "$O/ncu_launches.log" 2>&1
echo "launches_rc=$?"
ncu --set full --clock-control none --import-source on -k regex:masked_scan -s 4 -c 1 -o "$O/masked_scan_general" $B > /dev/null 2>&1
echo "scan_rc=$?"
ncu --set full --clock-control none --import-source on -k regex:merge_ln -s 4 -c 1 -o "$O/merge_ln" $B > /dev/null 2>&1
echo "merge_rc=$?"
i=16
for name in inproj xproj dtproj outproj; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s $i -c 1 -o "$O/gemm_tn_$name" $B > /dev/null 2>&1
  echo "gemm_${name}_rc=$?"
  i=$((i+1))
done
python tools/run_width.py 1280 50 > "$O/plain_1280.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"masked_scan|merge_ln" -s 6 -c 2 -o "$O/width1280_scan_merge" python tools/run_width.py 1280 50 > /dev/null 2>&1
echo "w1280_rc=$?"
python tools/run_width.py 640 50 > "$O/plain_640.log" 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"masked_scan|merge_ln" -s 6 -c 2 -o "$O/width640_scan_merge" python tools/run_width.py 640 50 > /dev/null 2>&1
echo "w640_rc=$?"
python tools/bench_configs.py > "$O/configs_3_5.jsonl" 2> "$O/configs.err"
echo "configs_rc=$?"
python tools/bench_latency.py > "$O/latency_graph.jsonl" 2>/dev/null
ls -la "$O"
