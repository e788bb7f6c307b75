#!/bin/bash
# usage: tools/r02_comm_sweep.sh <workload> "<z_chunks list>"   (2 GPUs; cfg4s8 reproduces the per-rank load of cfg4 on 8 GPUs)
WL=${1:-cfg4s8}; ZL=${2:-"1 2 4 8"}
for zc in $ZL; do
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$zc bench.py --gpus 2 --workload $WL --steps 10 --warmup 3 --no-e2e --no-same-workload --loop 0 --z-chunks $zc 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
k=c['rank0_kernel_times'] or {}
print('$WL z_chunks', $zc, 'ms', round(d['ms_per_step'],3), 'dec', round(c['dec_ms_rank0'],3), 'rec', round(c['rec_ms_rank0'],3), 'Mvox/s', round(d['value']), 'err', '%.2e'%c['dec_rel_err'], 'compute', round(k.get('compute_kernels_ms_per_step',0),2), 'exposed', round(k.get('exposed_comm_ms_per_step',0),2), 'timeouts', c['flag_wait_timeouts'])
"
done
