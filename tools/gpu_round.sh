#!/bin/bash
# one GPU visit: parity tests, bench, launch list, ncu --set full of the kernels named in $1 (regex), results under gpurun_out/<tag>_*
TAG=${2:-run}
KREGEX=${1:-k_chol}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
if [ -n "$KREGEX" ]; then
  ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" --launch-skip 40 -c 6 -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
cat gpurun_out/${TAG}_tests.log; head -c 1200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
