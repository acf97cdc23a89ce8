#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/tests_n2.log 2>&1
echo "exit $?" >> gpurun_out/tests_n2.log
tail -3 gpurun_out/tests_n2.log
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_n2_v10.log 2>&1
echo "exit $?" >> gpurun_out/bench_n2_v10.log
grep -E "^\{|^exit|rror" gpurun_out/bench_n2_v10.log | cut -c1-1500
