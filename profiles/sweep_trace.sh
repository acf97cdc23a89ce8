for cfg in "-DRTR_TRACE_MIN_CTAS=6 -DRTR_SMEM_STACK=12"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_render.py --force-build 2>&1 | tail -1
done
