#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_row_synthesis" 2>&1 | tail -4
for v in 0 2; do timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --loop 0 --rows-variant $v 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows_variant', $v, round(d['ms_per_step'],3), d['config']['pr_rel_err'], json.dumps({k:round(v['ms_per_launch'],3) for k,v in d['roofline']['all_kernels'].items()}))
"; done
