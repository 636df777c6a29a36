#!/bin/bash
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 900 python -m pytest tests/test_parity_mesh.py tests/test_streaming.py tests/test_edge_cases.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2 3; do
  timeout 600 python tools/bench_mesh.py 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['triangles'], round(d['extract_mesh_total_s'],3), round(d['serialize_data_s'],3), d['mesh_breakdown_ms'])"
done
