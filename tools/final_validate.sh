#!/bin/bash
# Round-end validation on one B200: whole GPU suite, smoke(), the default bench line and the other configs' lines,
# ncu launch list of the default bench command.
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -c 300 gpurun_out/final_smoke.log
timeout 400 python bench.py > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/r02_bench_cfg5.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_cfg5.json
for w in cfg3 cfg5haar; do
  timeout 300 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"
done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --loop 0"
$CMD > gpurun_out/r02_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_cfg5.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
