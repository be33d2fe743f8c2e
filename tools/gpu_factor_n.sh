#!/bin/bash
# factor-parallel series on one box: bench.py --mode factor for the 10k and C5 windows at every N given (subsets of the GPUs)
# usage: tools/gpu_factor_n.sh <tag> N [N ...]
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
for N in "$@"; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"
  for win in 10k C5; do
    NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 300 $TR bench.py --mode factor --window $win --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_factor_${win}_n$N.out 2> gpurun_out/${TAG}_factor_${win}_n$N.err
    grep -h "NCCL INFO.*nranks\|NVLS" gpurun_out/${TAG}_factor_${win}_n$N.err gpurun_out/${TAG}_factor_${win}_n$N.out | head -4 > gpurun_out/${TAG}_nccl_${win}_n$N.txt
    grep "^{" gpurun_out/${TAG}_factor_${win}_n$N.out > gpurun_out/${TAG}_factor_${win}_n$N.json; rm -f gpurun_out/${TAG}_factor_${win}_n$N.out
    python -c "
import json;l=json.load(open('gpurun_out/${TAG}_factor_${win}_n$N.json'));print('$win N=$N value',round(l['value']),'ms',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']))" || tail -3 gpurun_out/${TAG}_factor_${win}_n$N.err
  done
done
