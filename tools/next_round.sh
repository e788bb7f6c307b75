#!/bin/bash
# First GPU call of the next round: validate and time the variants prepared at the end of round 1
# (profiles/r01_scaling_and_experiments.md, "Prepared, not yet timed").  About 2 minutes on one B200:
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'tools/next_round.sh 2>&1 | tee gpurun_out/next_round.log'
set -u
for v in 600 700 800 2; do
  echo "== check NDDWT_VARIANT=$v (full-row synthesis variants)"
  NDDWT_ROWS_MIN_CTAS=0 NDDWT_VARIANT=$v timeout 60 python tools/check_variant.py 2>&1 | tail -2
done
echo "== check NDDWT_VARIANT=1000 (z-chunked full-row kernel for 3-D volumes)"
NDDWT_VARIANT=1000 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or 131x128x30 or 64x48x40" 2>&1 | tail -1
tools/variant_sweep.sh " " "0 2 600 700 800" cfg5
tools/variant_sweep.sh " " "0 1000" cfg3
