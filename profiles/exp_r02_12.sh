#!/bin/bash
# N GPUs: pipeline variants selected by environment
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
i=0
for envs in "RTR_BENCH_COMM_PRIORITY=0" "RTR_BENCH_COMM_PRIORITY=-1"; do
i=$((i+1))
tag=n${N}_e$i
env $envs RTR_BENCH_WATCHDOG=200 RTR_BENCH_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+i)) bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline $2 > gpurun_out/pipe_$tag.log 2> gpurun_out/pipe_$tag.err
echo "exit $?" >> gpurun_out/pipe_$tag.log
python - <<PY
import json
for l in open("gpurun_out/pipe_$tag.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$tag", "$envs", "value %.0f (%.2f ms)  e2e %.0f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["multi_gpu"]["stripes_of_rank"], d["image_check"]["ok"])
    elif l.startswith("exit"):
        print(l.strip())
PY
grep -E "rror|Traceback" gpurun_out/pipe_$tag.err | head -5
grep "^rank 0: " gpurun_out/pipe_$tag.err | head -1 | cut -c1-420
done
