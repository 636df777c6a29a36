#!/bin/bash
# round-2 evidence run (GPU box): tests, bench lines of both arms, launch list of the bench command
# (timed regions only), full captures of the dominant kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_tests.log
timeout 900 python bench.py > gpurun_out/bench_ours_r2.json 2> gpurun_out/bench_ours_r2.err; tail -c 1500 gpurun_out/bench_ours_r2.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; tail -c 600 gpurun_out/bench_ref_r2.json
MRH_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/launches_r2.json 2> gpurun_out/launches_r2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 30 -c 2 -o gpurun_out/r2_frame_final python tools/profile_frames.py 34 > gpurun_out/r2_ncu_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_blocks -c 1 -o gpurun_out/r2_mc_kernel python tools/bench_mesh.py 20000 0.005 >> gpurun_out/r2_ncu_final.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv
