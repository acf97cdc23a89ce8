#!/bin/bash
# one gpurun call: full GPU parity suite, then build / traversal sweeps
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
{
for cfg in "" "-DRTR_PLOC_MINB=8" "-DRTR_PLOC_WARPS=8 -DRTR_PLOC_MINB=3" "-DRTR_PLOC_WARPS=8 -DRTR_PLOC_MINB=4" "-DRTR_PLOC_WARPS=2 -DRTR_PLOC_MINB=12"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_build.py --force-build 2>&1 | tail -2 | cut -c1-400
done
for g in 32 64 128; do
  echo "RTR_L2_FETCH=$g"; RTR_L2_FETCH=$g timeout 300 python profiles/time_build.py 2>&1 | tail -2 | cut -c1-400
done
} > gpurun_out/sweep_ploc2.log 2>&1
cat gpurun_out/sweep_ploc2.log
{
for cfg in "" "-DRTR_TRACE_MIN_CTAS=7" "-DRTR_TRACE_MIN_CTAS=8" "-DRTR_SMEM_STACK=8 -DRTR_TRACE_MIN_CTAS=7" "-DRTR_SMEM_STACK=8 -DRTR_TRACE_MIN_CTAS=8" "-DRTR_SMEM_STACK=32"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_render.py --force-build 2>&1 | tail -1
done
for g in 32 64 128; do
  echo "RTR_L2_FETCH=$g"; RTR_L2_FETCH=$g timeout 300 python profiles/time_render.py 2>&1 | tail -1
done
} > gpurun_out/sweep_trace2.log 2>&1
cat gpurun_out/sweep_trace2.log
