#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_mesh.py tests/test_mesh_full_size.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2 3; do
  timeout 600 python tools/bench_mesh.py 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['triangles'], round(d['extract_mesh_total_s'],3), d['mesh_breakdown_ms'])"
done
MRH_BENCH_REF=1 timeout 600 python tools/bench_lidar.py 40 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())
for k,v in d.items():
    if isinstance(v,dict): print(k, round(v['frames_per_sec_e2e']), round(v['frames_per_sec_e2e_pipelined']), round(v.get('reference_frames_per_sec_e2e',0)), round(v['device_ms_per_frame'],3), v['host_us_per_frame'])"
