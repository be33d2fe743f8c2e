#!/bin/bash
# GPU visit for the large-window path: C5 parity test, C5 bench (window-parallel x256), launch list
TAG=${1:-c5}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --config C5 --no-cpu --steps 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --config C5 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
cat gpurun_out/${TAG}_tests.log; python tools/launch_table.py gpurun_out/${TAG}_launches.csv 2>/dev/null | head -12
python -c "
import json;l=json.load(open('gpurun_out/${TAG}_bench.json'));print('value',round(l['value']),'ms',l['ms_per_step'],'e2e',round(l['e2e']['value']),l['latency'], l['stage_share'])"
tail -3 gpurun_out/${TAG}_bench.err
