#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
( time python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/tests_gpu.log 2>&1
tail -6 gpurun_out/tests_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -4 gpurun_out/bench_default.err
cut -c1-400 gpurun_out/bench_default.json
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -4 gpurun_out/bench_reference.err
cut -c1-600 gpurun_out/bench_reference.json
