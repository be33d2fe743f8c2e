#!/bin/bash
# window-parallel bench at N ranks with sleeping (default when LOCAL_WORLD_SIZE > 1) and spinning host waits
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633"
for v in 1 0 1 0; do
  echo "UVS_BLOCKING_SYNC=$v N=$N"
  UVS_BLOCKING_SYNC=$v timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2))"
done
