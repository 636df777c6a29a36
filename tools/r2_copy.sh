#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_copy.log
python tools/h2d_probe.py 640 480 > $L 2>&1; python tools/h2d_probe.py 1280 960 >> $L 2>&1
for env in "" "MRH_ONE_COPY_STREAM=1"; do
  for rep in 1 2; do
    env $env timeout 600 python bench.py --steps 500 --warmup 30 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$env', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['host_us_per_step'], 'pageable', round(d['e2e_pageable']['value']))" >> $L
  done
done
timeout 600 python -m pytest tests/test_ingest_pipeline.py tests/test_parity_rgbd.py tests/test_streaming.py -m gpu -x -q 2>&1 | tail -2 >> $L
cat $L
