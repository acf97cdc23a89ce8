#!/bin/bash
# Round-2 evidence on one B200: GPU tests, the default bench line, the reference arm, the ncu launch list of the bench
# command, ncu --set full captures of the kernels DESIGN.md quotes.  Outputs under gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_tests_gpu.log; cat gpurun_out/r02_tests_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 400 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
D="python profiles/prof_driver.py --frames 2"
N="timeout 400 ncu --set full --clock-control none --import-source on"
$N -k regex:trace_persistent -s 1 -c 1 -o gpurun_out/prof_r02_render $D > gpurun_out/ncu_r02.log 2>&1
$N -k regex:ploc_loop -s 1 -c 1 -o gpurun_out/prof_r02_ploc_loop $D >> gpurun_out/ncu_r02.log 2>&1
$N -k regex:flatten_emit -s 1 -c 1 -o gpurun_out/prof_r02_flatten_emit $D >> gpurun_out/ncu_r02.log 2>&1
$N -k regex:flatten_positions -s 1 -c 1 -o gpurun_out/prof_r02_flatten_positions $D >> gpurun_out/ncu_r02.log 2>&1
$N -k regex:onesweep -s 4 -c 1 -o gpurun_out/prof_r02_onesweep $D >> gpurun_out/ncu_r02.log 2>&1
$N -k regex:leaf_init -s 1 -c 1 -o gpurun_out/prof_r02_leaf $D >> gpurun_out/ncu_r02.log 2>&1
ls -la gpurun_out/prof_r02_*.ncu-rep; tail -3 gpurun_out/ncu_r02.log
