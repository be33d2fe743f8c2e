#!/bin/bash
# quick GPU visit: chain-solve phase cycles (timing build, one window), single-window stage times, short bench + launch list
TAG=${1:-q}
mkdir -p gpurun_out
UVS_LIB=tools/probes/libuvs_timing.so timeout 120 python tools/latency_probe.py 1 2>&1 | grep -v "^  phase" | sort | uniq -c | sort -rn | head -8 > gpurun_out/${TAG}_timing.txt
timeout 120 python tools/latency_probe.py 1 > gpurun_out/${TAG}_lat.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q -k "full_solve or first_step or large_batch or rejected" 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --no-cpu --steps 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
cat gpurun_out/${TAG}_timing.txt gpurun_out/${TAG}_lat.txt gpurun_out/${TAG}_tests.log; python tools/launch_table.py gpurun_out/${TAG}_launches.csv | head -12; python -c "
import json;l=json.load(open('gpurun_out/${TAG}_bench.json'));print('value',round(l['value']),'e2e',round(l['e2e']['value']),l['latency'], l['stage_share'])"
tail -3 gpurun_out/${TAG}_bench.err
