#!/bin/bash
# N GPUs: the frame schedule with the traces
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
tag=n${N}_d
RTR_BENCH_WATCHDOG=200 RTR_BENCH_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29850 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/pipe_$tag.log 2> gpurun_out/pipe_$tag.err
echo "exit $?" >> gpurun_out/pipe_$tag.log
python - <<PY
import json
for l in open("gpurun_out/pipe_$tag.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$tag", "value %.0f (%.2f ms)  e2e %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["multi_gpu"], d["image_check"]["ok"])
    elif l.startswith("exit"):
        print(l.strip())
PY
grep -E "rror|Traceback" gpurun_out/pipe_$tag.err | head -5
grep "^rank [01]: " gpurun_out/pipe_$tag.err | cut -c1-520
