#!/bin/bash
# N GPUs (default 2): the re-ordered pipeline (broadcast of f+2 beside the rays of f), with and without the green-context partition
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
i=0
for extra in ""; do
i=$((i+1))
tag=n${N}_v$i
RTR_BENCH_WATCHDOG=240 RTR_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+i)) bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline $extra > gpurun_out/pipe_$tag.log 2> gpurun_out/pipe_$tag.err
echo "exit $?" >> gpurun_out/pipe_$tag.log
python - <<PY
import json
for l in open("gpurun_out/pipe_$tag.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$tag", "$extra", "value %.0f (%.2f ms)  e2e %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d.get("multi_gpu"), d["image_check"]["ok"])
    elif l.startswith("exit"):
        print(l.strip())
PY
grep -E "rror|Traceback" gpurun_out/pipe_$tag.err | head -5
done
