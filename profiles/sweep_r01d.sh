#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_sort_gpu.py tests/test_build_gpu.py tests/test_cpp_harness.py -x -q -m gpu > gpurun_out/tests_sort_build.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_sort_build.log
tail -4 gpurun_out/tests_sort_build.log
{
for cfg in "" "-DRTR_SORT_LOOKBATCH=2" "-DRTR_SORT_LOOKBATCH=8" "-DRTR_SORT_PREFETCH=0" "-DRTR_SORT_PREFETCH=592" "-DRTR_SORT_BLOCK=256 -DRTR_SORT_IPT=16 -DRTR_SORT_MINB=4 -DRTR_SORT_PREFETCH=592" "-DRTR_SORT_BLOCK=256 -DRTR_SORT_IPT=12 -DRTR_SORT_MINB=4 -DRTR_SORT_PREFETCH=592" "-DRTR_SORT_BLOCK=384 -DRTR_SORT_IPT=16 -DRTR_SORT_MINB=2"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_sort.py --force-build 2>&1 | tail -1
done
} > gpurun_out/sweep_sort2.log 2>&1
cat gpurun_out/sweep_sort2.log
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 300 python profiles/time_build.py 2>&1 | tail -2 | cut -c1-400 | tee gpurun_out/time_build_d.log
D="python profiles/prof_driver.py --frames 2 --no-render"
N="ncu --set full --clock-control none --import-source on"
$N -k regex:onesweep -s 4 -c 1 -o gpurun_out/prof_d_onesweep $D > gpurun_out/ncu_d.log 2>&1
$N -k regex:ploc_iteration -s 48 -c 3 -o gpurun_out/prof_d_ploc $D >> gpurun_out/ncu_d.log 2>&1
$N -k regex:flatten_level -s 57 -c 1 -o gpurun_out/prof_d_flatten $D >> gpurun_out/ncu_d.log 2>&1
