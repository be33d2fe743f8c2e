#!/bin/bash
# compute-sanitizer over the paths touched in round 2: memcheck on the smoke run (every kernel of a small window), on the
# large-window solve, the IMU / prior sweeps and the marginalization; racecheck (shared-memory hazards) on the smoke run
TAG=${1:-san}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${TAG}_memcheck_smoke.txt 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/${TAG}_memcheck_smoke.txt
UVS_FUSE_MIN=1 timeout 900 $S --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${TAG}_memcheck_smoke_fused.txt 2>&1; echo "memcheck smoke (fused path) rc=$?"; tail -2 gpurun_out/${TAG}_memcheck_smoke_fused.txt
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "large_window or every_window_of_a_batch or factor_sweep_parity or prior_parity or relocalisation" > gpurun_out/${TAG}_memcheck_tests.txt 2>&1; echo "memcheck tests rc=$?"; tail -4 gpurun_out/${TAG}_memcheck_tests.txt
timeout 900 $S --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${TAG}_racecheck_smoke.txt 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/${TAG}_racecheck_smoke.txt
