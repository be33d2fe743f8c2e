#!/bin/bash
# ncu --set full of the materialised sweep kernels (one launch each) + plain timing; results under gpurun_out/<tag>_*
TAG=${1:-sw}
mkdir -p gpurun_out
python tools/sweep_probe.py 10 > gpurun_out/${TAG}_sweep.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_line_vp|k_imu_geom|k_imu_weight|k_proj|k_prior$" --launch-skip 10 -c 5 -f -o gpurun_out/${TAG}_sweep python tools/sweep_probe.py 3 > gpurun_out/${TAG}_sweep_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_sweep.ncu-rep --page raw --csv > gpurun_out/${TAG}_sweep_raw.csv 2>/dev/null
cat gpurun_out/${TAG}_sweep.txt; tail -3 gpurun_out/${TAG}_sweep_ncu.log
