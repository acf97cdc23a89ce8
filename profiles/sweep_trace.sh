# parameter sweep of the persistent traversal (10 M triangles, 4K, 2 bounces); macros: RTR_TRACE_MIN_CTAS, RTR_SMEM_STACK,
# RTR_WALK_STEPS (walk steps per bookkeeping round), RTR_BLOCK_BATCH / RTR_LEAF_BATCH (when the leaf step runs), RTR_PREFETCH
for cfg in "" "-DRTR_WALK_STEPS=2" "-DRTR_BLOCK_BATCH=2" "-DRTR_LEAF_BATCH=16" "-DRTR_TRACE_MIN_CTAS=7" "-DRTR_PREFETCH=0"; do
  RTR_NVCC_EXTRA="$cfg" timeout 300 python profiles/time_render.py --force-build 2>&1 | tail -1
done
