#!/bin/bash
mkdir -p gpurun_out
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
export NCCL_DEBUG=WARN
i=0
for cfg in "4 4" "8 8" "2 2" "0 16"; do
  set -- $cfg
  i=$((i+1))
  if [ "$1" != "0" ]; then export NCCL_MAX_NCHANNELS=$1; export NCCL_MIN_NCHANNELS=1; else unset NCCL_MAX_NCHANNELS; fi
  RTR_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus 2 --steps 8 --warmup 3 --no-extras --no-cpu-baseline --reserve-sms $2 > gpurun_out/bench_n2_$i.log 2>&1
  echo "cfg channels=$1 reserve=$2: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2_$i.log | head -1) $(grep -o '"broadcast_ms": [0-9.]*' gpurun_out/bench_n2_$i.log)"
  grep "^rank 1" gpurun_out/bench_n2_$i.log | head -1 | cut -c1-400
done
