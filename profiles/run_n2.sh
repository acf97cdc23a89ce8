#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/tests_n2.log 2>&1
echo "exit $?" >> gpurun_out/tests_n2.log
tail -4 gpurun_out/tests_n2.log
export NCCL_DEBUG=WARN
RTR_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1
echo "exit $?" >> gpurun_out/bench_n2.log
grep -E "^\{|^exit|rror" gpurun_out/bench_n2.log | cut -c1-2500
grep "^rank 0" gpurun_out/bench_n2.log | head -1 | cut -c1-700
