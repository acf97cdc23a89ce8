#!/bin/bash
# the rebuild's CTA shapes again, now that the PLOC loop is one persistent launch
for cfg in "" "-DRTR_PLOC_WARPS=4 -DRTR_PLOC_MINB=8" "-DRTR_PLOC_WARPS=4 -DRTR_PLOC_MINB=6" "-DRTR_PLOC_WARPS=16 -DRTR_PLOC_MINB=2" "-DRTR_PLOC_WARPS=8 -DRTR_PLOC_MINB=3"; do
  RTR_BUILD_ONLY=ploc.cu RTR_NVCC_EXTRA="$cfg" timeout 200 python profiles/time_build.py --force-build 2>&1 | tail -2 | cut -c1-330
done
