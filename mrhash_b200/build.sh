#!/bin/bash
# Builds libmrhash_b200.so in-tree for sm_100a (called by __graft_entry__.build()).
set -e
cd "$(dirname "$0")/csrc"
NVCC=${NVCC:-nvcc}
FLAGS="-std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC,-O2,-Wall,-fopenmp"
mkdir -p ../_build
pids=()
for f in mrh_capi mrh_frame mrh_state mrh_mesh mrh_weld mrh_halo mrh_grid; do
  [ -f $f.cu ] || continue
  if [ ! -f ../_build/$f.o ] || [ -n "$(find . ../../include -newer ../_build/$f.o \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    $NVCC $FLAGS -c $f.cu -o ../_build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ../libmrhash_b200.so ../_build/*.o -lcudart -lgomp
echo "built $(realpath ../libmrhash_b200.so)"
