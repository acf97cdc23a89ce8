RTR_BUILD_ONLY=sort.cu RTR_NVCC_EXTRA=-DRTR_SORT_PHASE_CLOCKS python -m realtimeraytracing_b200.build --force 2>&1 | grep -i error
python profiles/sort_phases.py
