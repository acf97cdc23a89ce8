# ncu --set full captures of the kernels DESIGN.md quotes (one GPU; read the .ncu-rep files with profiles/ncu_summary.py
# and profiles/ncu_lines.py).  run_bench_r01.sh does the same after the bench and its launch list.
set -x
D="python profiles/prof_driver.py --frames 2"
N="ncu --set full --clock-control none --import-source on"
$N -k regex:onesweep -s 4 -c 1 -o gpurun_out/prof_onesweep $D > gpurun_out/ncu.log 2>&1
$N -k regex:ploc_iteration -s 48 -c 2 -o gpurun_out/prof_ploc $D >> gpurun_out/ncu.log 2>&1
$N -k regex:leaf_init -s 1 -c 1 -o gpurun_out/prof_leaf $D >> gpurun_out/ncu.log 2>&1
$N -k regex:flatten_emit -s 1 -c 1 -o gpurun_out/prof_flatten_emit $D >> gpurun_out/ncu.log 2>&1
$N -k regex:trace_persistent -s 1 -c 1 -o gpurun_out/prof_render $D >> gpurun_out/ncu.log 2>&1
