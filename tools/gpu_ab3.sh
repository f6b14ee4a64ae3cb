#!/bin/bash
# Timing-only A/B of library variants on C4 (whole frame, timed split, part of 8) and C3.
# usage: bash tools/gpu_ab3.sh variant[:ENV=VAL] ...
for combo in "$@"; do
  IFS=: read v e <<< "$combo"
  lib=$PWD/eraytracer_b200/lib/libert_b200.so; [ "$v" != base ] && lib=$PWD/eraytracer_b200/lib/libert_b200_$v.so
  echo "== $combo"
  env ERT_B200_LIB=$lib ${e:-X=1} python tools/part_probe.py c4 1 8 2>&1 | head -2
  env ERT_B200_LIB=$lib ${e:-X=1} python tools/part_probe.py c4 8 8 2>&1 | head -1
  env ERT_B200_LIB=$lib ${e:-X=1} python tools/part_probe.py c3 1 8 2>&1 | head -1
done
