#!/bin/bash
# Builds variants of one translation unit (extra -D flags) into most_queue_b200/_variants/lib_<name>.so for A/B runs
# on the same GPU box:   tools/ab_variants.sh k3_onepass.cu base "" lean "-DOP_V_X=1" ...
#   then on the box:     MQB200_LIB=most_queue_b200/_variants/lib_lean.so python tools/bench_scan.py 28
set -e
cd "$(dirname "$0")/.."
SRC=$1; shift
mkdir -p most_queue_b200/_variants
OBJ=most_queue_b200/csrc/_build
while [ $# -gt 0 ]; do
  name=$1; defs=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 186 $defs \
       -c most_queue_b200/csrc/$SRC -o most_queue_b200/_variants/${SRC%.cu}_$name.o
  others=$(ls $OBJ/*.o | grep -v "/${SRC%.cu}.o")
  nvcc -shared -o most_queue_b200/_variants/lib_$name.so $others most_queue_b200/_variants/${SRC%.cu}_$name.o -gencode arch=compute_100a,code=sm_100a
  echo built lib_$name.so
done
