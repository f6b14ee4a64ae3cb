#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -4 gpurun_out/s3_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/s3_$name.json 2> gpurun_out/s3_$name.err || tail -5 gpurun_out/s3_$name.err; }
ARGS="--workload c4" run c4_mb3 A=1
ARGS="--workload c4" run c4_mb2 ERT_B200_LIB=$PWD/eraytracer_b200/lib/libert_b200_mb2.so
ARGS="--workload c4" run c4_mb4 ERT_B200_LIB=$PWD/eraytracer_b200/lib/libert_b200_mb4.so
ARGS="--workload c3" run c3_mb3 A=1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s3_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f.split('/')[-1], "ms %.2f Mrays/s %.0f frac %.4f launches %d e2e_ms %.2f exact %.3g box %.3g filt %.3g"%(d["ms_per_step"], d["value"], r["frac"], d["gpu_launches"], d["e2e"]["ms_per_step"], r["exact_fp64_sphere_tests"], r["box_tests"], r["sphere_filter_tests"]))
    except Exception as e: print(f, "fail", e)
PY
