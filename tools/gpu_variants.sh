#!/bin/bash
# A/B of library build variants on one workload: tools/gpu_variants.sh <workload> <variant>...
mkdir -p gpurun_out
W=$1; shift
for v in "$@"; do
  ERT_B200_LIB=$PWD/eraytracer_b200/lib/libert_b200_$v.so timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_${W}_$v.json 2> gpurun_out/var_${W}_$v.err || tail -5 gpurun_out/var_${W}_$v.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/var_${W}_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f.split('/')[-1], "ms %.2f Mrays/s %.0f frac %.4f launches %d exact %.3g box %.3g filt %.3g"%(d["ms_per_step"], d["value"], r["frac"], d["gpu_launches"], r["exact_fp64_sphere_tests"], r["box_tests"], r["sphere_filter_tests"]))
    except Exception as e: print(f, "fail", e)
PY
