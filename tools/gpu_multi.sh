#!/bin/bash
# multi-GPU visit on N GPUs of one box: 2-GPU parity test of the factor-parallel mode, bench.py --mode factor (10k and C5
# windows) and the window-parallel bench line at N ranks; results under gpurun_out/<tag>_*
# usage: tools/gpu_multi.sh <tag> <N>
TAG=${1:-m}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
for win in 10k C5; do
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 $TR bench.py --mode factor --window $win --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_factor_${win}_n$N.json 2> gpurun_out/${TAG}_factor_${win}_n$N.err
  grep -h "NCCL INFO.*nranks\|NVLS\|comm 0x.*rank" gpurun_out/${TAG}_factor_${win}_n$N.err gpurun_out/${TAG}_factor_${win}_n$N.json | head -6 > gpurun_out/${TAG}_nccl_${win}_n$N.txt
  grep "^{" gpurun_out/${TAG}_factor_${win}_n$N.json > gpurun_out/${TAG}_factor_${win}_n$N.line; mv gpurun_out/${TAG}_factor_${win}_n$N.line gpurun_out/${TAG}_factor_${win}_n$N.json
done
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2> gpurun_out/${TAG}_window_n$N.err | grep "^{" > gpurun_out/${TAG}_window_n$N.json
cat gpurun_out/${TAG}_tests.log 2>/dev/null
for f in gpurun_out/${TAG}_factor_*_n$N.json gpurun_out/${TAG}_window_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.load(open(sys.argv[1])); print(sys.argv[1], 'value', round(l['value']), 'ms/step', round(l['ms_per_step'],3), 'e2e', round(l['e2e']['value']), l['config'].get('parallelism'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
tail -3 gpurun_out/${TAG}_factor_10k_n$N.err | cut -c1-300
