for cfg in "-DRTR_SORT_IPT=16" "-DRTR_SORT_IPT=12" "-DRTR_SORT_IPT=8" "-DRTR_SORT_IPT=10"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_sort.py --force-build 2>&1 | tail -1
done
