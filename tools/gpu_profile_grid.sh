#!/bin/bash
# full ncu captures of the path kernels of the cell-grid wavefront on C4 (optional: ERT_B200_LIB variant)
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --accel grid --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/grid_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_trace_path -s 5 -c 5 -o gpurun_out/prof_grid $CMD > gpurun_out/grid_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
