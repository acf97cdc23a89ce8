#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_tests_gpu.log; cat gpurun_out/r02_tests_gpu.log
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/cub_sort_baseline.cu -o /tmp/cub_sort_baseline && /tmp/cub_sort_baseline | tee gpurun_out/r02_cub_sort_baseline.txt
python profiles/sort_time.py 2>&1 | tail -6 | tee -a gpurun_out/r02_cub_sort_baseline.txt
