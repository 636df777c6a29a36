#!/bin/bash
# tuning sweep of the fused frame kernel (run on the GPU box)
mkdir -p gpurun_out
L=gpurun_out/r2_sweep1.log
: > $L
timeout 300 python -m pytest tests/test_parity_rgbd.py tests/test_edge_cases.py tests/test_fastdiv.py -m gpu -x -q >> $L 2>&1
for pref in 0/4 1/4 2/4 4/4; do MRH_FUSED_PREF=$pref timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L; done
MRH_BULK_DEPTH=0 timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
MRH_NO_FAST_DIV=1 timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
for c in 4 6; do MRH_FUSED_CTAS_PER_SM=$c timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L; done
MRH_FRAME=split timeout 120 python tools/bench_quick.py 200 2>&1 | grep flushed >> $L
timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed >> $L
MRH_FRAME=split timeout 200 python tools/bench_quick.py 150 1280 960 2000 2>&1 | grep flushed >> $L
cat $L
