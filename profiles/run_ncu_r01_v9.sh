#!/bin/bash
# ncu evidence for the final kernels of round 1 (v9): launch list of the bench command + --set full captures
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build > gpurun_out/build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01_v9.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r01_v9.csv
D="python profiles/prof_driver.py --frames 2"
N="timeout 300 ncu --set full --clock-control none --import-source on"
$N -k regex:trace_persistent -s 1 -c 1 -o gpurun_out/prof_v9_render $D > gpurun_out/ncu_v9.log 2>&1
$N -k regex:ploc_iteration -s 48 -c 2 -o gpurun_out/prof_v9_ploc $D >> gpurun_out/ncu_v9.log 2>&1
$N -k regex:onesweep -s 4 -c 1 -o gpurun_out/prof_v9_onesweep $D >> gpurun_out/ncu_v9.log 2>&1
$N -k regex:flatten_emit -s 1 -c 1 -o gpurun_out/prof_v9_flatten_emit $D >> gpurun_out/ncu_v9.log 2>&1
$N -k regex:pack_quads -s 1 -c 1 -o gpurun_out/prof_v9_pack_quads $D >> gpurun_out/ncu_v9.log 2>&1
ls -la gpurun_out/*v9*.ncu-rep
