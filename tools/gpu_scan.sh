#!/bin/bash
# brute-force scan A/B on C3: tools/gpu_scan.sh variant...
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "linear or scan or wavefront" 2>&1 | tail -2
for v in "$@"; do
  lib=$PWD/eraytracer_b200/lib/libert_b200.so; [ "$v" != base ] && lib=$PWD/eraytracer_b200/lib/libert_b200_$v.so
  ERT_B200_LIB=$lib timeout 300 python bench.py --workload c3 --accel linear --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/scan_$v.json 2> gpurun_out/scan_$v.err || tail -3 gpurun_out/scan_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/scan_$v.json')); r=d['roofline']; print('$v', 'ms %.1f'%d['ms_per_step'], r['frame']['class_frac'], r['frame']['ms'])"
done
