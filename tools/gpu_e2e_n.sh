#!/bin/bash
# window-parallel bench at N ranks of one box, twice (device-resident value and end-to-end time per step)
# usage: tools/gpu_e2e_n.sh [N]     (profiles/r2_e2e_n8_sync_ab.txt was taken with this script while a sleeping-wait variant of
#                                    the library existed: UVS_BLOCKING_SYNC, see profiles/r2_notes.md)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633"
for rep in 1 2; do
  echo "N=$N run $rep"
  timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2))"
done
