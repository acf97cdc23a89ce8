#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests/test_trace_gpu.py -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
{
for cfg in "" "-DRTR_LEAF_BATCH=14"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_render.py --force-build 2>&1 | tail -1
done
} > gpurun_out/sweep_trace3.log 2>&1
cat gpurun_out/sweep_trace3.log
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 300 python profiles/time_build.py 2>&1 | tail -2 | cut -c1-500 | tee gpurun_out/time_build_f.log
