#!/bin/bash
# Experiment builds: every translation unit to its own object (parallel, cached by mtime), then one link.
#   scripts/expbuild.sh out.so [-DNAME=VALUE ...]     extra flags apply to ALL units compiled in this call
# Objects live in /tmp/cs_objs/<flags-hash>/; the product build stays cudasift_b200/build.py.
set -e
cd "$(dirname "$0")/../cudasift_b200/csrc"
OUT=$1; shift
TAG=$(echo "$*" | md5sum | cut -c1-8)
OBJ=/tmp/cs_objs/$TAG; mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -I../../include $*"
pids=()
for f in api pipeline2 pyramid pyramid2 detect detect2 cap32 describe match match_tc homography geom; do
  if [ ! -f $OBJ/$f.o ] || [ $f.cu -nt $OBJ/$f.o ] || [ common.cuh -nt $OBJ/$f.o ] || [ tma.cuh -nt $OBJ/$f.o ] || [ pipeline2.h -nt $OBJ/$f.o ]; then
    nvcc $FLAGS -c $f.cu -o $OBJ/$f.o & pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared $OBJ/*.o -o "$OUT"
echo "built $OUT"
