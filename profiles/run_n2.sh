#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_trace_gpu.py -x -q -m gpu > gpurun_out/tests_n2.log 2>&1
echo "exit $?" >> gpurun_out/tests_n2.log
tail -25 gpurun_out/tests_n2.log
