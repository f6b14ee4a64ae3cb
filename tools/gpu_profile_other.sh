#!/bin/bash
# full ncu captures of the shading / emission / shadow kernels of one C4 frame
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"wf_shade|wf_emit_hits|wf_trace_shadow|wf_finalize" -s 8 -c 9 -o gpurun_out/prof_other $CMD > gpurun_out/other_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
