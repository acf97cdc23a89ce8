#!/bin/bash
# On the GPU box: times every prebuilt variant (profiles/make_variants.sh) with the given timing script.
#   bash profiles/sweep_variants.sh profiles/time_render.py [args...]
LIB=realtimeraytracing_b200/lib
cp $LIB/librtr_b200.so /tmp/librtr_b200.keep
for v in $LIB/variants/librtr_b200.*.so; do
  name=$(basename $v .so); name=${name#librtr_b200.}
  cp $v $LIB/librtr_b200.so; touch $LIB/librtr_b200.so
  echo -n "$name: "; timeout 300 python "$@" 2>&1 | tail -1
done
cp /tmp/librtr_b200.keep $LIB/librtr_b200.so; touch $LIB/librtr_b200.so
