#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
export NCCL_DEBUG=WARN
RTR_BENCH_TRACE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --reserve-sms 24 > gpurun_out/bench_n8.log 2>&1
echo "exit $?" >> gpurun_out/bench_n8.log
grep -E "^\{|^exit|rror" gpurun_out/bench_n8.log | cut -c1-2600; grep "^rank [01]:" gpurun_out/bench_n8.log | head -2 | cut -c1-900
