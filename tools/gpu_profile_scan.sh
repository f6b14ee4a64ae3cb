#!/bin/bash
# full ncu capture of the brute-force scan kernels on C3 (--accel linear)
mkdir -p gpurun_out
CMD="python bench.py --workload c3 --accel linear --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:wf_scan -s 4 -c 3 -o gpurun_out/prof_scan $CMD > gpurun_out/scan_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
