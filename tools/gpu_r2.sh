#!/bin/bash
# one GPU visit (round 2): parity tests, bench, ncu launch list; results under gpurun_out/<tag>_*
# usage: tools/gpu_r2.sh <tag> [pytest -k expression] [ncu --set full kernel regex]
TAG=${1:-r2}
KEXPR=${2:-}
KREGEX=${3:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
else
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
fi
timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
if [ -n "$KREGEX" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" --launch-skip 30 -c 6 -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
cat gpurun_out/${TAG}_tests.log; head -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
