#!/bin/bash
# one GPU, e2e arm: host copies beside the whole frame (default) or starting with the rays
mkdir -p gpurun_out
for v in "" 1; do
RTR_BENCH_E2E_COPIES_WITH_RAYS=$v timeout 200 python bench.py --no-cpu-baseline --no-extras > gpurun_out/e2e_v$v.json 2> gpurun_out/e2e_v$v.err
python - <<PY
import json
for l in open("gpurun_out/e2e_v$v.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("copies_with_rays='$v': value %.1f (%.2f ms)  e2e %.1f (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["image_check"]["ok"])
PY
tail -c 200 gpurun_out/e2e_v$v.err
done
