#!/bin/bash
mkdir -p gpurun_out
for v in "" _mc32; do
  for rep in 1 2; do
    MRH_LIB=$PWD/mrhash_b200/libmrhash_b200$v.so timeout 600 python tools/bench_mesh.py 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['triangles'], d['extract_mesh_total_s'], d['mesh_breakdown_ms'])"
  done
done 2>&1 | tee gpurun_out/r2_mc_sweep.log
