#!/bin/bash
# one GPU visit recording the state of the tree: parity tests, bench lines of every BASELINE config, launch list
# usage: tools/gpu_state.sh <tag>
TAG=${1:-s}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 20 --check 8 > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_C2_reference.json 2>> gpurun_out/${TAG}_bench_C2.err
for c in C1 C5 10k; do
  timeout 600 python bench.py --config $c --steps 5 --check 2 --cpu-budget 6 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
for c in C5 10k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_$c.log 2>&1
done
cat gpurun_out/${TAG}_tests.log; for c in C2 C1 C5 10k; do head -c 600 gpurun_out/${TAG}_bench_$c.json; echo; tail -3 gpurun_out/${TAG}_bench_$c.err; done
