for cfg in "-DRTR_PLOC_PP=2" "-DRTR_PLOC_PP=4" "-DRTR_PLOC_PP=1"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_build.py --force-build 2>&1 | tail -2 | cut -c1-330
done
