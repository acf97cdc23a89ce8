#!/bin/bash
# N=2: rebuild / broadcast / rays against each other
mkdir -p gpurun_out
for r in 16; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29611+r)) profiles/overlap_n2.py $r 4 > gpurun_out/overlap_n2_r$r.log 2>&1
echo "exit $?" >> gpurun_out/overlap_n2_r$r.log
grep -E "^\[rank|^exit|rror" gpurun_out/overlap_n2_r$r.log | cut -c1-400
done
