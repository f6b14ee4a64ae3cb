#!/bin/bash
# bench lines for the workloads + ncu launch list + one full capture of the top kernel
mkdir -p gpurun_out
for w in c4 c3 c2 c2d5; do
  python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || echo "bench $w failed"
  tail -c 600 gpurun_out/bench_$w.err
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:render_free -s 4 -c 1 -o gpurun_out/prof_c4 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
