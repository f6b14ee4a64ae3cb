#!/bin/bash
# ncu --set full of the path kernels of one part of 8 of the C4 frame
mkdir -p gpurun_out
python tools/part_probe.py c4 8 2 > gpurun_out/part_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wf_trace_path|wf_trace_shadow|wf_shade" -s 60 -c 14 -f -o gpurun_out/prof_part8 python tools/part_probe.py c4 8 2 > gpurun_out/part_ncu.log 2>&1
tail -2 gpurun_out/part_plain.log; tail -2 gpurun_out/part_ncu.log
