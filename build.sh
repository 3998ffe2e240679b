#!/bin/bash
# Builds libzoo_sm100.so in-tree for sm_100a (the only architecture this library targets).
set -e
cd "$(dirname "$0")"
PKG="reinforcementlearningzoo.jl_b200"
SRC="$PKG/csrc"
OUT="$PKG/libzoo_sm100.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v"
mkdir -p build
objs=""
pids=""
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  objs="$objs $o"
  if [ ! -f $o ] || [ $f -nt $o ] || [ $SRC/common.cuh -nt $o ] || [ include/zoo_sm100.h -nt $o ] || ls $SRC/*.cuh | while read h; do [ $h -nt $o ] && echo y; done | grep -q y; then
    ( $NVCC $FLAGS -c $f -o $o > build/$(basename ${f%.cu}).log 2>&1 || { cat build/$(basename ${f%.cu}).log; exit 1; } ) &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $objs -lcudart
echo "built $OUT"
