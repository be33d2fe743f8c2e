#!/bin/bash
# end-of-round records: bench lines of every BASELINE config (with the CPU baseline and a sampled parity check of the timed
# batch), the reference arm, the N = 1 points of the factor-parallel series, launch lists, one --set full capture of the
# top solver kernels.  Results under gpurun_out/<tag>_*.
TAG=${1:-fin}
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.txt 2>&1; tail -1 gpurun_out/${TAG}_smoke.txt
# ncu --set full of the materialised sweep kernels first: the bench lines below report its DRAM bytes as roofline.traffic
bash tools/gpu_sweep_ncu.sh ${TAG} > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_sweep.ncu-rep gpurun_out/${TAG}_sweep_ncu.json --windows 1184 --sweep --command "ncu --set full --clock-control none --import-source on -k regex:k_line_vp|k_imu_geom|k_imu_weight|k_proj|k_prior\$ --launch-skip 10 -c 5 python tools/sweep_probe.py 3" && cp gpurun_out/${TAG}_sweep_ncu.json profiles/r2_sweep_ncu.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_C2_reference.json 2> gpurun_out/${TAG}_ref.err
timeout 900 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err
timeout 900 python bench.py --check 8 --no-cpu --steps 5 > gpurun_out/${TAG}_bench_C2_check.json 2> gpurun_out/${TAG}_bench_C2.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_C2.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
for c in C1 C5 10k; do
  timeout 900 python bench.py --config $c --steps 5 --cpu-budget 6 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/${TAG}_launches_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
done
bash tools/gpu_factor1.sh ${TAG}
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_chol_chain|k_lin_lines|k_lin_points|k_window_system|k_window_tail|k_back" --launch-skip 40 -c 7 -f -o gpurun_out/${TAG}_solver python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_solver.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_solver.ncu-rep gpurun_out/${TAG}_solver_ncu.json --windows 1184 --command "ncu --set full --clock-control none --import-source on -k regex:k_chol_chain|k_lin_lines|k_lin_points|k_window_system|k_window_tail|k_back --launch-skip 40 -c 7 python bench.py --steps 1 --warmup 1 --no-cpu"
for c in C2 C1 C5 10k; do python -c "
import json;l=json.load(open('gpurun_out/${TAG}_bench_$c.json'));print('$c value',round(l['value']),'ms',round(l['ms_per_step'],2),'e2e',round(l['e2e']['value']),'cpu',l['cpu_baseline'] and (round(l['cpu_baseline']['value']), round(l['cpu_baseline']['single_thread_value']), l['cpu_baseline']['cores']), 'lat', round(l['latency']['single_window_ms_per_solve'],3), l.get('parity_check'))"; done
cat gpurun_out/${TAG}_bench_C2_reference.json | cut -c1-600
