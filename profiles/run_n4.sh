#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
export NCCL_DEBUG=WARN
for n in 2 4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_n$n.log 2>&1
echo "exit $?" >> gpurun_out/bench_n$n.log
grep -E "^\{|^exit|rror" gpurun_out/bench_n$n.log | cut -c1-230
grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_n$n.log; grep -o '"multi_gpu": {[^}]*}' gpurun_out/bench_n$n.log
done
