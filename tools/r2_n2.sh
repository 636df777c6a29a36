#!/bin/bash
mkdir -p gpurun_out
python tools/h2d_probe.py 640 480 > gpurun_out/r2_h2d.log 2>&1; python tools/h2d_probe.py 1280 960 >> gpurun_out/r2_h2d.log 2>&1; cat gpurun_out/r2_h2d.log
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_n2_tests.log; cat gpurun_out/r2_n2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 30 > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err; tail -c 3000 gpurun_out/bench_n2_r2.json; tail -5 gpurun_out/bench_n2_r2.err
timeout 600 python bench.py --width 1280 --height 960 --steps 500 --warmup 30 --no-extras --no-cpu-baseline > gpurun_out/bench_1280x960_r2.json 2> gpurun_out/bench_1280x960_r2.err; tail -c 1500 gpurun_out/bench_1280x960_r2.json
