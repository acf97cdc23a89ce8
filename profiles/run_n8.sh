#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
export NCCL_DEBUG=WARN
for r in 8 16; do
RTR_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29570+r)) bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --reserve-sms $r > gpurun_out/bench_n8_r$r.log 2>&1
echo "exit $?" >> gpurun_out/bench_n8_r$r.log
grep -E "^\{|^exit|rror" gpurun_out/bench_n8_r$r.log | cut -c1-230
grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_n8_r$r.log; grep -o '"multi_gpu": {[^}]*}' gpurun_out/bench_n8_r$r.log
grep "^rank 1" gpurun_out/bench_n8_r$r.log | head -1 | cut -c1-420
done
