#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_sweep3.log
: > $L
timeout 400 python -m pytest tests/test_parity_rgbd.py tests/test_edge_cases.py tests/test_fastdiv.py -m gpu -x -q >> $L 2>&1
timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
for v in c6; do MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_$v.so timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L; done
for pf in 0/4 2/4 4/4; do MRH_FUSED_PREF=$pf timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L; done
MRH_BULK_DEPTH=0 timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_dbg.so timeout 200 python tools/debug_fused.py 12 2>&1 | tail -7 >> $L
timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed >> $L
MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_c6.so timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed >> $L
cat $L
