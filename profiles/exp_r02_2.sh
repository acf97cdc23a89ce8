timeout 600 python -m pytest tests/test_sort_gpu.py -m gpu -x -q 2>&1 | tail -3
python profiles/sort_time.py 10 2>&1 | tail -4
for cfg in "1 6" "1 12" "1 2" "0 4"; do set -- $cfg; echo "SORT RANK_ATOMIC=$1 GROUP=$2"; RTR_BUILD_ONLY=sort.cu RTR_NVCC_EXTRA="-DRTR_SORT_RANK_ATOMIC=$1 -DRTR_SORT_RANK_GROUP=$2" python -m realtimeraytracing_b200.build --force 2>&1 | grep -i "error\|spill" | head -3; python profiles/sort_time.py 10 2>&1 | head -2; done
