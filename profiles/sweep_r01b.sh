#!/bin/bash
# one gpurun call: parity tests of the rewritten sort / PLOC kernels, then parameter sweeps
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_sort_gpu.py tests/test_build_gpu.py -x -q -m gpu > gpurun_out/tests_sort_build.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_sort_build.log
tail -5 gpurun_out/tests_sort_build.log
{
for cfg in "" "-DRTR_SORT_IPT=16" "-DRTR_SORT_IPT=8" "-DRTR_SORT_BLOCK=256 -DRTR_SORT_IPT=16 -DRTR_SORT_MINB=4" "-DRTR_SORT_BLOCK=384 -DRTR_SORT_IPT=16 -DRTR_SORT_MINB=2" "-DRTR_SORT_BLOCK=256 -DRTR_SORT_IPT=20 -DRTR_SORT_MINB=3" "-DRTR_SORT_LOOKBATCH=8" "-DRTR_SORT_LOOKBATCH=4"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_sort.py --force-build 2>&1 | tail -1
done
} > gpurun_out/sweep_sort.log 2>&1
cat gpurun_out/sweep_sort.log
{
for cfg in "" "-DRTR_PLOC_ROLL=1" "-DRTR_PLOC_MINB=5" "-DRTR_PLOC_MINB=8" "-DRTR_PLOC_WARPS=8 -DRTR_PLOC_MINB=3" "-DRTR_PLOC_WARPS=8 -DRTR_PLOC_MINB=3 -DRTR_PLOC_ROLL=1" "-DRTR_PLOC_WARPS=2 -DRTR_PLOC_MINB=12"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_build.py --force-build 2>&1 | tail -2 | cut -c1-400
done
} > gpurun_out/sweep_ploc.log 2>&1
cat gpurun_out/sweep_ploc.log
