#!/bin/bash
TAG=${1:-im}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_imu_geom|k_imu_weight" --launch-skip 4 -c 2 -f -o gpurun_out/${TAG}_imu python tools/sweep_probe.py 3 > gpurun_out/${TAG}_imu_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_imu.ncu-rep --page raw --csv > gpurun_out/${TAG}_imu_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_imu.ncu-rep --page source --csv --kernel-name k_imu_weight > gpurun_out/${TAG}_imu_src.csv 2>/dev/null
tail -3 gpurun_out/${TAG}_imu_ncu.log
