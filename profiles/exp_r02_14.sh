#!/bin/bash
# sweep of the drain's threshold: a worker's 8/57 of a 4-bounce frame and the full frames
mkdir -p gpurun_out
for v in 2 8 16; do
RTR_BUILD_ONLY=trace.cu RTR_NVCC_EXTRA="-DRTR_STEAL_MIN_IDLE=$v" python -m realtimeraytracing_b200.build --force > /dev/null 2>&1
echo "== RTR_STEAL_MIN_IDLE=$v"
python profiles/trace_time.py 2 3 2>&1 | tail -3
done
RTR_BUILD_ONLY=trace.cu RTR_NVCC_EXTRA="-DRTR_STEAL=0" python -m realtimeraytracing_b200.build --force > /dev/null 2>&1
echo "== RTR_STEAL=0"
python profiles/trace_time.py 2 3 2>&1 | tail -3
