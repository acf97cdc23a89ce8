set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sb in 0 2 4 8 16; do echo "TILE_BLOCK=$sb"; RTR_TILE_BLOCK=$sb python profiles/trace_time.py 2 5 2>&1 | tail -1; done
python profiles/sort_time.py 10 2>&1 | tail -4
for cfg in "256 16 4" "256 12 4" "512 8 3" "384 16 3" "512 16 2" "256 24 3"; do set -- $cfg; echo "SORT BLOCK=$1 IPT=$2 MINB=$3"; RTR_BUILD_ONLY=sort.cu RTR_NVCC_EXTRA="-DRTR_SORT_BLOCK=$1 -DRTR_SORT_IPT=$2 -DRTR_SORT_MINB=$3" python -m realtimeraytracing_b200.build --force 2>&1 | grep -i "error\|spill" | head -3; python profiles/sort_time.py 10 2>&1 | head -2; done
