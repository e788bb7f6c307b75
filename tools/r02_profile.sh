#!/bin/bash
# Round-2 bench lines of the other BASELINE configs + ncu launch list and full capture of the two tile kernels (cfg5).
for w in cfg3 cfg1 cfg2 cfg5haar; do
  timeout 400 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  echo "$w rc=$?"; tail -c 200 gpurun_out/r02_bench_$w.json
done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --loop 0"
$CMD > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_cfg5.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_dec3_fused|k_rec3_rows" -s 8 -c 2 -o gpurun_out/r02_prof_cfg5 -f $CMD > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"
