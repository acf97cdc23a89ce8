#!/bin/bash
# final evidence of round 1: smoke, pytest -m gpu, default bench, reference arm, launch list and ncu capture of the traversal
bash profiles/run_final_r01.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01_v10.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r01_v10.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 1 -c 1 -o gpurun_out/prof_v10_render python profiles/prof_driver.py --frames 2 > gpurun_out/ncu_v10.log 2>&1
ls -la gpurun_out/prof_v10_render.ncu-rep
