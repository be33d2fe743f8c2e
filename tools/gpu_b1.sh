#!/bin/bash
# GPU visit for the single-window path: parity tests, solve latency at B = 1 (and 16), launch list of one single-window solve
TAG=${1:-b1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_tests.log
timeout 120 python tools/latency_probe.py 1 > gpurun_out/${TAG}_lat.txt 2>&1
timeout 120 python tools/latency_probe.py 16 >> gpurun_out/${TAG}_lat.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_lat_launches.csv python tools/latency_probe.py 1 > /dev/null 2>&1
cat gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_lat.txt; python tools/launch_table.py gpurun_out/${TAG}_lat_launches.csv 2>/dev/null | head -24
