# parameter sweep of the Onesweep sort (10 M pairs / keys); macros: RTR_SORT_BLOCK, RTR_SORT_IPT, RTR_SORT_MINB,
# RTR_SORT_LOOKBATCH (predecessors per look-back round), RTR_SORT_PREFETCH (tiles ahead pulled into L2)
for cfg in "" "-DRTR_SORT_LOOKBATCH=2" "-DRTR_SORT_LOOKBATCH=8" "-DRTR_SORT_PREFETCH=0" "-DRTR_SORT_BLOCK=256 -DRTR_SORT_IPT=16 -DRTR_SORT_MINB=4 -DRTR_SORT_PREFETCH=592"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_sort.py --force-build 2>&1 | tail -1
done
