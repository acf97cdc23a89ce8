#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trace_gpu.py -m gpu -q -k "drain" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_tests_multi_gpu.log
