#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_sweep9.log
: > $L
for v in "" _to1 _to2; do
  echo "variant $v" >> $L
  MRH_LIB=$PWD/mrhash_b200/libmrhash_b200$v.so timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
  MRH_LIB=$PWD/mrhash_b200/libmrhash_b200$v.so timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed=True >> $L
done
cat $L
