#!/bin/bash
# Builds tuning variants of the library: meyda_b200/_lib/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
OUT=meyda_b200/_lib/variants; mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2"
build() { # name, extra flags
  name=$1; shift
  objs=""
  for f in capi kernel_generic kernel_warp kernel_warp_mf kernel_exact_warp; do nvcc $FLAGS "$@" -c meyda_b200/csrc/$f.cu -o $OUT/${name}_$f.o -Xptxas -v 2> $OUT/${name}_$f.ptxas.txt; objs="$objs $OUT/${name}_$f.o"; done
  nvcc -shared -o $OUT/lib_$name.so $objs -lcudart_static -lpthread -ldl -lrt
  grep -A2 "warp2048_kernelILj262143ELb0" $OUT/${name}_kernel_warp.ptxas.txt | grep -E "registers|spill" | tr '\n' ' '; echo " <- $name"
}
for v in "$@"; do case $v in [0-9]*) build w$v -DMB_WARPS=$v;; *+*) build $v $(echo "-D$v" | sed "s/+/ -D/g");; *) build $v -D$v;; esac; done
