#!/bin/bash
# Builds variants of librtr_b200.so HERE (nvcc cross-compiles) that differ in the -D flags of ONE source file, so that a
# sweep on the GPU box only swaps libraries:   bash profiles/make_variants.sh trace.cu name1 "-DA=1" name2 "-DB=2" ...
# -> realtimeraytracing_b200/lib/variants/librtr_b200.<name>.so   (git-ignored, travels with gpurun)
set -e
cd "$(dirname "$0")/.."
SRC=$1; shift
python -m realtimeraytracing_b200.build > /dev/null 2>&1
LIB=realtimeraytracing_b200/lib
mkdir -p $LIB/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-fvisibility=default,-ffp-contract=off --expt-relaxed-constexpr -ccbin /usr/bin/g++"
pids=()
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  (
    nvcc $FLAGS $defs -c realtimeraytracing_b200/csrc/$SRC -o $LIB/variants/${SRC%.cu}.$name.o
    objs=""
    for o in $LIB/obj/*.o; do
      if [ "$(basename $o)" = "${SRC%.cu}.o" ]; then objs="$objs $LIB/variants/${SRC%.cu}.$name.o"; else objs="$objs $o"; fi
    done
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o $LIB/variants/librtr_b200.$name.so $objs -lcudart_static -ldl -lpthread -lrt
    rm -f $LIB/variants/${SRC%.cu}.$name.o
    echo "built $name [$defs]"
  ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
ls -la $LIB/variants/
