#!/bin/bash
# Sweep NDDWT_VARIANT tuning variants: parity subset first, then short benches (no CPU baseline / e2e).
# usage: tools/variant_sweep.sh "<variants for parity>" "<variants for bench>" [workloads]
PV=${1:-"0"}; BV=${2:-"0"}; WL=${3:-"cfg5"}
for v in $PV; do
  echo "== parity NDDWT_VARIANT=$v"
  NDDWT_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or parity_vs_oracle or fused_equals_generic" 2>&1 | tail -2
done
for w in $WL; do for v in $BV; do
  NDDWT_VARIANT=$v timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    k=d['roofline']['all_kernels']
    print('$w v=$v', 'Mvox/s', round(d['value'],1), 'dec', round(d['config']['dec_ms'],3), 'rec', round(d['config']['rec_ms'],3), 'pair_frac', round(d['roofline']['pair_frac'],3), 'err', '%.2e'%d['config']['pr_rel_err'], ' '.join('%s=%.3f'%(n.split()[-1],x['ms_per_launch']) for n,x in k.items()))
except Exception as e:
    print('$w v=$v FAILED', e)
"
done; done
