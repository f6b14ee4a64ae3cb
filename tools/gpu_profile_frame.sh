#!/bin/bash
# ncu --set full of the kernels of ONE C4 frame matching a regex: bash tools/gpu_profile_frame.sh <name> <regex> [skip] [count]
mkdir -p gpurun_out
python tools/frame_once.py c4 3 > gpurun_out/frame_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-0} -c ${4:-6} -f -o gpurun_out/prof_$1 python tools/frame_once.py c4 1 > gpurun_out/frame_ncu_$1.log 2>&1
tail -2 gpurun_out/frame_plain.log; tail -3 gpurun_out/frame_ncu_$1.log
