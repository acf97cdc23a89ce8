set -x
python -m realtimeraytracing_b200.build --force > gpurun_out/build.log 2>&1
D="python profiles/prof_driver.py --frames 2 --no-render"
N="ncu --set full --clock-control none --import-source on"
$N -k regex:onesweep -s 4 -c 2 -o gpurun_out/prof_b_onesweep $D > gpurun_out/ncu_b.log 2>&1
$N -k regex:ploc_iteration -s 48 -c 6 -o gpurun_out/prof_b_ploc $D >> gpurun_out/ncu_b.log 2>&1
grep -c "PROF" gpurun_out/ncu_b.log
ls -la gpurun_out
