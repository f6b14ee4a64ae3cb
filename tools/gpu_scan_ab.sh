#!/bin/bash
# A/B of scan-kernel build variants and ERT_SCAN_WAVES on the C4 and C3 bands (tools/scan_band.py)
mkdir -p gpurun_out; : > gpurun_out/scan_ab.txt
for combo in "$@"; do
  IFS=: read v waves kind <<< "$combo"
  lib=$PWD/eraytracer_b200/lib/libert_b200.so; [ "$v" != base ] && lib=$PWD/eraytracer_b200/lib/libert_b200_$v.so
  echo "== $combo" >> gpurun_out/scan_ab.txt
  ERT_B200_LIB=$lib ERT_SCAN_WAVES=$waves timeout 300 python tools/scan_band.py ${kind:-c4} 3 2>&1 | tail -2 >> gpurun_out/scan_ab.txt
done
cat gpurun_out/scan_ab.txt
