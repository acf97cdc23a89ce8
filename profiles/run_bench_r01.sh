#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
echo "exit $?" >> gpurun_out/bench_r01.err
cut -c1-3000 gpurun_out/bench_r01.json
# launch list of the same command (cold-cache, serialised per-launch times: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r01.csv
# full captures of the kernels the summaries quote
D="python profiles/prof_driver.py --frames 2"
N="ncu --set full --clock-control none --import-source on"
$N -k regex:onesweep -s 4 -c 1 -o gpurun_out/prof_r01_onesweep $D > gpurun_out/ncu_r01.log 2>&1
$N -k regex:ploc_iteration -s 48 -c 2 -o gpurun_out/prof_r01_ploc $D >> gpurun_out/ncu_r01.log 2>&1
$N -k regex:leaf_init -s 1 -c 1 -o gpurun_out/prof_r01_leaf $D >> gpurun_out/ncu_r01.log 2>&1
$N -k regex:flatten_emit -s 1 -c 1 -o gpurun_out/prof_r01_flatten_emit $D >> gpurun_out/ncu_r01.log 2>&1
$N -k regex:flatten_level -s 29 -c 1 -o gpurun_out/prof_r01_flatten_level $D >> gpurun_out/ncu_r01.log 2>&1
$N -k regex:trace_persistent -s 1 -c 1 -o gpurun_out/prof_r01_render $D >> gpurun_out/ncu_r01.log 2>&1
ls -la gpurun_out/*.ncu-rep
