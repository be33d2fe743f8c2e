#!/bin/bash
# GPU visit: parity tests, then the materialised sweep with the occupancy variants of k_line_vp
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for v in 3 4 5; do echo "UVS_LINE_OCC=$v"; UVS_LINE_OCC=$v python tools/sweep_probe.py 20; done 2>&1 | tee gpurun_out/${TAG}_sweep_ab.txt
