#!/bin/bash
# Round-2 closing evidence on one B200 (after the traversal drain): GPU tests, smoke, the default bench line, the reference
# arm, the ncu launch list of the bench command, one ncu --set full capture of the traversal kernel, compute-sanitizer.
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_tests_gpu.log; cat gpurun_out/r02_tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 1 -c 1 -o gpurun_out/prof_r02b_render python profiles/prof_driver.py --frames 2 > gpurun_out/ncu_r02b.log 2>&1
ls -la gpurun_out/prof_r02b_render.ncu-rep; tail -2 gpurun_out/ncu_r02b.log
SAN_TIMEOUT=200 bash profiles/sanitize.sh
