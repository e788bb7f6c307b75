#!/bin/bash
# Sweep the 1-D cascade variants of a tuning build (NDDWT_CASC): parity subset first, then short cfg2 benches.
# usage: tools/casc_sweep.sh "<variants for parity>" "<variants for bench>"
PV=${1:-"0"}; BV=${2:-"0"}
for v in $PV; do
  echo "== parity NDDWT_CASC=$v"
  NDDWT_CASC=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shrink.py -m gpu -x -q -k "${KSEL:-1d or cfg2 or batched or golden or parity_vs_oracle or shrink}" 2>&1 | tail -3
done
for v in $BV; do
  NDDWT_CASC=$v timeout 200 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('cfg2 casc=$v', 'Msamples/s', round(d['value'],1), 'dec', round(d['config']['dec_ms'],3), 'rec', round(d['config']['rec_ms'],3), 'pair_frac', round(d['roofline']['pair_frac'],3), 'err', '%.2e'%d['config']['pr_rel_err'])
except Exception as e:
    print('cfg2 casc=$v FAILED', e)
"
done
