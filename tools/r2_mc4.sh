#!/bin/bash
for rep in 1 2 3; do
for v in "" _noroll; do
  MRH_LIB=$PWD/mrhash_b200/libmrhash_b200$v.so timeout 600 python tools/bench_mesh.py 40000 0.004 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['blocks'], d['triangles'], round(d['extract_mesh_total_s'],3), {k:round(x,1) for k,x in d['mesh_breakdown_ms'].items()})"
done
done
