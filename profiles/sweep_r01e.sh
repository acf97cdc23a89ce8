#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 300 python profiles/time_build.py 2>&1 | tail -2 | cut -c1-500 | tee gpurun_out/time_build_e.log
