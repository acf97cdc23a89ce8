#!/bin/bash
mkdir -p gpurun_out
RTR_BUILD_ONLY=trace.cu RTR_NVCC_EXTRA="-DRTR_STEAL_STATS $1" python -m realtimeraytracing_b200.build --force > /dev/null 2>&1
python profiles/steal_stats.py 4 2>&1 | tee gpurun_out/steal_stats.log | tail -16
