#!/bin/bash
# standard GPU visit of round 2: parity tests, single-window latency, C2 bench, launch list, sweep ncu capture (-> profiles/r2_sweep_ncu.json)
TAG=${1:-v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_tests.log
timeout 120 python tools/latency_probe.py 1 > gpurun_out/${TAG}_lat.txt 2>&1
timeout 600 python bench.py --no-cpu --steps 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
bash tools/gpu_sweep_ncu.sh ${TAG} > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_sweep.ncu-rep gpurun_out/${TAG}_sweep_ncu.json --windows 1184 --sweep --command "ncu --set full --clock-control none --import-source on -k regex:k_line_vp|k_imu_geom|k_imu_weight|k_proj|k_prior\$ --launch-skip 10 -c 5 python tools/sweep_probe.py 3"
cat gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_lat.txt; python tools/launch_table.py gpurun_out/${TAG}_launches.csv 2>/dev/null | head -16
python -c "
import json;l=json.load(open('gpurun_out/${TAG}_bench.json'));print('value',round(l['value']),'e2e',round(l['e2e']['value']),l['latency'], l['stage_share'], l['roofline']['frac'], l['roofline']['traffic'])"
tail -3 gpurun_out/${TAG}_bench.err
