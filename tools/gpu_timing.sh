#!/bin/bash
# phase cycles of the chain solve from the -DUVS_CHOL_TIMING build: one window alone, then window 0 of a full batch
for B in 1 1184; do
  echo "== B=$B"
  UVS_LIB=tools/probes/libuvs_timing.so timeout 120 python tools/latency_probe.py $B 2>&1 | grep -v "^B=\|^stages" | tail -6
done
