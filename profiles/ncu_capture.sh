set -x
D="python profiles/prof_driver.py --frames 2"
N="ncu --set full --clock-control none --import-source on"
$N -k regex:onesweep -s 4 -c 1 -o gpurun_out/prof_s2_onesweep $D > gpurun_out/ncu_s2.log 2>&1
$N -k regex:ploc_iteration -s 48 -c 3 -o gpurun_out/prof_s2_ploc $D >> gpurun_out/ncu_s2.log 2>&1
$N -k regex:leaf_init -s 1 -c 1 -o gpurun_out/prof_s2_leaf $D >> gpurun_out/ncu_s2.log 2>&1
$N -k regex:flatten_level -s 57 -c 1 -o gpurun_out/prof_s2_flatten $D >> gpurun_out/ncu_s2.log 2>&1
$N -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_s2_render $D >> gpurun_out/ncu_s2.log 2>&1
grep -c "PROF" gpurun_out/ncu_s2.log
