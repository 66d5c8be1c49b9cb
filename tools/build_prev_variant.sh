#!/bin/bash
# Builds the committed (HEAD) kernels as meyda_b200/_lib/variants/lib_prev.so for A/B runs against the working tree.
set -e
cd "$(dirname "$0")/.."
OUT=meyda_b200/_lib/variants; mkdir -p $OUT; TMP=$(mktemp -d)
mkdir -p $TMP/meyda_b200/csrc $TMP/include
git show ${1:-HEAD}:include/meyda_b200.h > $TMP/include/meyda_b200.h
for f in $(git ls-tree --name-only ${1:-HEAD} meyda_b200/csrc/ | xargs -n1 basename); do git show ${1:-HEAD}:meyda_b200/csrc/$f > $TMP/meyda_b200/csrc/$f; done
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2"
OBJS=""
for f in $(ls $TMP/meyda_b200/csrc/*.cu | xargs -n1 basename | sed 's/\.cu$//'); do nvcc $FLAGS -c $TMP/meyda_b200/csrc/$f.cu -o $OUT/prev_$f.o; OBJS="$OBJS $OUT/prev_$f.o"; done
nvcc -shared -o $OUT/lib_prev.so $OBJS -lcudart_static -lpthread -ldl -lrt 2>/dev/null
rm -rf $TMP; ls -la $OUT/lib_prev.so
