run() { echo "=== $1"; shift; env "$@" RTR_BENCH_WATCHDOG=50 timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras $EXTRA > gpurun_out/x8.json 2> gpurun_out/x8.err; grep "^\[rank 0\]" gpurun_out/x8.err | tail -2; tail -1 gpurun_out/x8.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['image_check']['ok'], d['multi_gpu'])
except Exception as e: print('no json', e)"; }
EXTRA="--no-partition" run "N=8 no partition" A=1
EXTRA="" run "N=8 partition, two ray streams" A=1
