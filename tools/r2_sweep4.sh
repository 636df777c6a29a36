#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_sweep4.log
: > $L
timeout 400 python -m pytest tests/test_parity_rgbd.py tests/test_edge_cases.py -m gpu -x -q >> $L 2>&1
timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
for pf in 0/4 1/8 2/4; do MRH_FUSED_PREF=$pf timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L; done
MRH_LIB=$PWD/mrhash_b200/libmrhash_b200_dbg.so timeout 200 python tools/debug_fused.py 40 2>&1 | tail -12 >> $L
timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed >> $L
cat $L
