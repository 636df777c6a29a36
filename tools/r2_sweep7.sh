#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_sweep7.log
: > $L
timeout 600 python -m pytest tests -m gpu -x -q >> $L 2>&1
timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
for v in c6; do MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_$v.so timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L; done
MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_dbg.so timeout 200 python tools/debug_fused.py 31 2>&1 | tail -22 | head -12 >> $L
timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed >> $L
cat $L
