#!/bin/bash
# one GPU: parity of the traversal kernel after a change + the frame time
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trace_gpu.py tests/test_configs_gpu.py -m gpu -x -q > gpurun_out/exp9_tests.log 2>&1
echo "exit $?" >> gpurun_out/exp9_tests.log
tail -4 gpurun_out/exp9_tests.log
for b in 2 4; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --bounces $b > gpurun_out/exp9_bench_b$b.json 2> gpurun_out/exp9_bench_b$b.err
python - <<PY
import json
for l in open("gpurun_out/exp9_bench_b$b.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("bounces $b: value %.1f (%.2f ms)  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["image_check"]["ok"], d["image_check"]["hash"])
PY
tail -2 gpurun_out/exp9_bench_b$b.err
done
python profiles/trace_time.py 2>&1 | tail -6; RTR_BUILD_ONLY=1 true
