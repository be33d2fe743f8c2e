#!/bin/bash
# chain-solve time against batch size for both variants (stage timer 'chol', us per iteration)
for pm in 0 100000; do
for B in 1 16 64 148 296 592; do
  echo -n "pipe_max $pm B $B: "
  UVS_CHAIN_PIPE_MAX=$pm timeout 120 python tools/latency_probe.py $B 2>&1 | grep "^stages\|profiling=0" | sed -e "s/.*'chol': \([0-9.]*\).*/chol \1 us/" | tr '\n' ' '; echo
done
done
timeout 300 python -m pytest tests -m gpu -x -q -k "full_solve or first_step or large_batch or rejected" 2>&1 | tail -2
