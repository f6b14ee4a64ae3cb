#!/bin/bash
# ncu --set full on the brute-force scan kernel over the C4 band (path launches of bounce 0 and 2, one shadow launch)
mkdir -p gpurun_out
K=${1:-c4}
python tools/scan_band.py $K 2 > gpurun_out/scan_band_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_scan -c 4 -f -o gpurun_out/prof_scan_$K python tools/scan_band.py $K 1 > gpurun_out/scan_ncu_full.log 2>&1
tail -3 gpurun_out/scan_band_plain.log; tail -5 gpurun_out/scan_ncu_full.log
