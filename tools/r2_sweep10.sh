#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_sweep10.log
: > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
for v in "" _hf; do
  echo "variant $v" >> $L
  MRH_LIB=$PWD/mrhash_b200/libmrhash_b200$v.so timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
  MRH_LIB=$PWD/mrhash_b200/libmrhash_b200$v.so timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed=True >> $L
done
MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_dbg.so timeout 200 python tools/debug_fused.py 31 2>&1 | grep -v "^cta" | tail -11 >> $L
MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_hfdbg.so timeout 200 python tools/debug_fused.py 31 2>&1 | grep -v "^cta" | tail -11 >> $L
cat $L
