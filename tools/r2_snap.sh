#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ingest_pipeline.py tests/test_parity_rgbd.py tests/test_parity_lidar.py -m gpu -x -q 2>&1 | tail -2
python tools/h2d_probe.py 640 480 | head -2
for rep in 1 2; do
  timeout 600 python bench.py --steps 600 --warmup 30 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v,1) for k,v in d['e2e']['host_us_per_step'].items()}, 'pageable', round(d['e2e_pageable']['value']))"
done
python tools/debug_timeline.py 2>&1 | tail -5
