#!/bin/bash
# factor mode on one GPU (the N = 1 point of the strong-scaling series)
TAG=${1:-m}
for win in 10k C5; do
  timeout 600 python bench.py --mode factor --window $win --steps 10 --warmup 3 2> gpurun_out/${TAG}_factor_${win}_n1.err | grep "^{" > gpurun_out/${TAG}_factor_${win}_n1.json
  python -c "
import json;l=json.load(open('gpurun_out/${TAG}_factor_${win}_n1.json'));print('$win n1 value',round(l['value']),'ms',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']))"
done
