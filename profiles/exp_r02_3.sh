run() { echo "=== $1"; shift; env "$@" RTR_BENCH_WATCHDOG=35 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-extras $EXTRA > gpurun_out/x.json 2> gpurun_out/x.err; grep "^\[rank\|File \"/root/repo" gpurun_out/x.err | tail -3; tail -1 gpurun_out/x.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['image_check']['ok'], d['multi_gpu']['stripes_of_rank'])
except Exception as e: print('no json', e)"; }
EXTRA="" run "partition, two ray streams" A=1
