#!/bin/bash
# Tuning build: tools/build_variant.sh <name> <extra nvcc flags...>  ->  mrhash_b200/libmrhash_b200_<name>.so
# (use with MRH_LIB=... ; the product build is mrhash_b200/build.sh)
set -e -o pipefail
cd "$(dirname "$0")/../mrhash_b200/csrc"
NAME=$1; shift
mkdir -p ../_build_$NAME
for f in mrh_capi mrh_frame mrh_state mrh_mesh mrh_weld mrh_halo mrh_grid; do
  nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC,-O2,-fopenmp "$@" -c $f.cu -o ../_build_$NAME/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ../libmrhash_b200_$NAME.so ../_build_$NAME/*.o -lcudart -lgomp
echo "built libmrhash_b200_$NAME.so"
