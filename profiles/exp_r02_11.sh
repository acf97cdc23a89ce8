#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trace_gpu.py tests/test_configs_gpu.py -m gpu -x -q > gpurun_out/exp11_tests.log 2>&1
echo "exit $?" >> gpurun_out/exp11_tests.log
tail -4 gpurun_out/exp11_tests.log
python profiles/trace_time.py 2 2>&1 | tail -3
python profiles/trace_time.py 4 2>&1 | tail -3 | head -1
bash profiles/exp_r02_10.sh "$1"
