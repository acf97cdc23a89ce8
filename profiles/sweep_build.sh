# parameter sweep of the rebuild (per-stage device times, 10 M triangles); macros: RTR_PLOC_WARPS (warps per CTA of
# ploc_iteration_kernel), RTR_PLOC_MINB (resident CTAs the register budget is cut for), RTR_PLOC_ROLL
for cfg in "" "-DRTR_PLOC_WARPS=4 -DRTR_PLOC_MINB=8" "-DRTR_PLOC_WARPS=8 -DRTR_PLOC_MINB=3" "-DRTR_PLOC_ROLL=1"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_build.py --force-build 2>&1 | tail -2 | cut -c1-400
done
