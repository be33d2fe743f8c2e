#!/bin/bash
# round-end evidence: bench line (with the CPU baseline), ncu launch list of the same command, ncu --set full of the sweep
# kernels and of the heaviest solver kernels; everything lands in gpurun_out/<tag>_*
TAG=${1:-final}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
UVS_SERIAL=1 ncu --set full --clock-control none --import-source on -k "regex:k_proj|k_line_vp|k_imu_geom|k_imu_weight|k_prior$" --launch-skip 60 -c 12 -f -o gpurun_out/${TAG}_sweep python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_sweep.log 2>&1
UVS_SERIAL=1 ncu --set full --clock-control none --import-source on -k "regex:k_chol|k_window_system|k_direct_fused|k_window_tail|k_core_points|k_core_lines|k_back_lines|k_back_points|k_step" --launch-skip 90 -c 9 -f -o gpurun_out/${TAG}_solver python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_solver.log 2>&1
head -c 1500 gpurun_out/${TAG}_bench.json; echo; head -c 600 gpurun_out/${TAG}_bench_reference.json; echo; tail -3 gpurun_out/${TAG}_bench.err
