#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards): smoke run on both landmark paths, a 48-window batch (barrier-phased chain solve)
TAG=${1:-rc}
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${TAG}_racecheck_smoke.txt 2>&1; echo "racecheck smoke rc=$?"; tail -2 gpurun_out/${TAG}_racecheck_smoke.txt
UVS_FUSE_MIN=1 timeout 900 $S --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${TAG}_racecheck_smoke_fused.txt 2>&1; echo "racecheck smoke (fused) rc=$?"; tail -2 gpurun_out/${TAG}_racecheck_smoke_fused.txt
timeout 1500 $S --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "large_batch_concurrent or large_window" > gpurun_out/${TAG}_racecheck_batch.txt 2>&1; echo "racecheck batch tests rc=$?"; tail -3 gpurun_out/${TAG}_racecheck_batch.txt
for f in gpurun_out/${TAG}_racecheck_*.txt; do echo "== $f"; grep -A3 "Race reported\|Error:" $f | grep -E " at | in " | sed -E 's/\+0x[0-9a-f]+//; s/=========\s+//' | sort | uniq -c | sort -rn | head -8; done
