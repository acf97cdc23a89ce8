set -x
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
D="python profiles/prof_driver.py --frames 1"
N="ncu --set full --clock-control none --import-source on"
$N -k regex:trace_persistent -c 1 -o gpurun_out/prof_g_render $D > gpurun_out/ncu_g.log 2>&1
tail -3 gpurun_out/ncu_g.log
