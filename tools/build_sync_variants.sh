#!/bin/bash
# A/B builds of kernel_warp.cu variants from given source files: name:source[:flags...]
cd "$(dirname "$0")/.."
OUT=meyda_b200/_lib/variants; mkdir -p meyda_b200/_lib/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2"
nvcc $FLAGS -c meyda_b200/csrc/capi.cu -o meyda_b200/_lib/variants/capi.o 2>/dev/null
nvcc $FLAGS -c meyda_b200/csrc/kernel_generic.cu -o meyda_b200/_lib/variants/kg.o 2>/dev/null
for spec in "$@"; do
  IFS=: read -r name src flags <<< "$spec"
  cp "$src" meyda_b200/csrc/_kw_tmp.cu
  nvcc $FLAGS $flags -c meyda_b200/csrc/_kw_tmp.cu -o meyda_b200/_lib/variants/kw_$name.o -Xptxas -v 2> meyda_b200/_lib/variants/kw_$name.txt
  nvcc -shared -o meyda_b200/_lib/variants/lib_$name.so meyda_b200/_lib/variants/capi.o meyda_b200/_lib/variants/kg.o meyda_b200/_lib/variants/kw_$name.o -lcudart_static -lpthread -ldl -lrt 2>/dev/null
  grep -A2 "kernelILb1ELb0" meyda_b200/_lib/variants/kw_$name.txt | grep -E "spill" | tr '\n' ' '; echo "<- $name"
done
rm -f meyda_b200/csrc/_kw_tmp.cu meyda_b200/_lib/variants/kw_*.o meyda_b200/_lib/variants/kw_*.txt meyda_b200/_lib/variants/capi.o meyda_b200/_lib/variants/kg.o
ls meyda_b200/_lib/variants
