#!/bin/bash
# A/B of library variants on the C4 frame: parity tests on the base build, then whole-frame and part-of-8 kernel times.
# usage: bash tools/gpu_ab2.sh [variant[:ENV=VAL] ...]   (variant = base or the suffix of libert_b200_<variant>.so)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_cell_grid.py tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_pytest.log
tail -4 gpurun_out/check_pytest.log
for combo in base "$@"; do
  IFS=: read v e <<< "$combo"
  lib=$PWD/eraytracer_b200/lib/libert_b200.so; [ "$v" != base ] && lib=$PWD/eraytracer_b200/lib/libert_b200_$v.so
  echo "== $combo"
  env ERT_B200_LIB=$lib ${e:-X=1} python tools/part_probe.py c4 1 8 2>&1 | head -2
  env ERT_B200_LIB=$lib ${e:-X=1} python tools/part_probe.py c4 8 8 2>&1 | head -1
  env ERT_B200_LIB=$lib ${e:-X=1} python tools/part_probe.py c3 1 8 2>&1 | head -1
done
