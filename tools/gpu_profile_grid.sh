#!/bin/bash
# ncu launch list (one steady-state frame) + full captures of the path kernels of the cell-grid wavefront on C4
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --accel grid --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/grid_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 44 --csv --log-file gpurun_out/grid_launches.csv $CMD > gpurun_out/grid_ncu_launches.log 2>&1
$CMD > gpurun_out/grid_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_trace_path -s 5 -c 5 -o gpurun_out/prof_grid $CMD > gpurun_out/grid_ncu_full.log 2>&1
ls -la gpurun_out | tail -4
