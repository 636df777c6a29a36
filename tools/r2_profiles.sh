#!/bin/bash
# round-2 evidence run (GPU box): launch list of the bench command, full captures of the dominant kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/launches_r2.json 2> gpurun_out/launches_r2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 30 -c 2 -o gpurun_out/r2_frame_final python tools/profile_frames.py 34 > gpurun_out/r2_ncu_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_points_emit|k_points_apply|k_alloc_points|Onesweep" -s 40 -c 6 -o gpurun_out/r2_lidar_kernels python tools/bench_lidar.py 6 2 >> gpurun_out/r2_ncu_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_blocks -c 1 -o gpurun_out/r2_mc_kernel python tools/bench_mesh.py 20000 0.005 >> gpurun_out/r2_ncu_final.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv
