#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_cpp_harness.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python profiles/time_render.py 2>&1 | tail -1
RTR_TILE_LPT=0 timeout 300 python profiles/time_render.py 2>&1 | tail -1
timeout 300 python profiles/time_render.py --width 1920 --height 1080 2>&1 | tail -1
RTR_TILE_LPT=0 timeout 300 python profiles/time_render.py --width 1920 --height 1080 2>&1 | tail -1
