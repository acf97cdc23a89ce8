#!/bin/bash
# last check of the committed tree on a fresh box: the driver's own sequence (GPU tests, smoke, default bench line)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err; cut -c1-220 gpurun_out/last_bench.json; tail -c 200 gpurun_out/last_bench.err
