#!/bin/bash
# ncu --set full of the kernels matching $2 during one bench step; report gpurun_out/<tag>_full.ncu-rep + csv pages
TAG=${1:-n}
KREGEX=${2:-k_lin}
SKIP=${3:-8}
CNT=${4:-4}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" --launch-skip $SKIP -c $CNT -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
tail -3 gpurun_out/${TAG}_ncu_full.log
