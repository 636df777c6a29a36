#!/bin/bash
# GPU box: mesh parity + 50 M-voxel mesh bench (in place vs forced round trip)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_mesh.py tests/test_streaming.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_mesh_tests3.log
cat gpurun_out/r2_mesh_tests3.log
timeout 600 python tools/bench_mesh.py > gpurun_out/r2_mesh_bench2.json 2> gpurun_out/r2_mesh_bench2.err; tail -1 gpurun_out/r2_mesh_bench2.json
MRH_MESH_ROUND_TRIP=1 timeout 600 python tools/bench_mesh.py > gpurun_out/r2_mesh_bench2_rt.json 2>> gpurun_out/r2_mesh_bench2.err; tail -1 gpurun_out/r2_mesh_bench2_rt.json
