#!/bin/bash
# parity tests + C4/C3 timings (A/B: hit binning on/off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -4 gpurun_out/s2_pytest.log
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/s2_$name.json 2> gpurun_out/s2_$name.err || tail -5 gpurun_out/s2_$name.err
}
ARGS="--workload c4" run c4_sort A=1
ARGS="--workload c4" run c4_nosort ERT_WF_NO_SORT=1
ARGS="--workload c3" run c3_sort A=1
ARGS="--workload c3" run c3_nosort ERT_WF_NO_SORT=1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s2_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; print(f.split('/')[-1], "ms %.2f Mrays/s %.0f frac %.4f launches %d e2e_ms %.2f"%(d["ms_per_step"], d["value"], r["frac"], d["gpu_launches"], d["e2e"]["ms_per_step"]))
    except Exception as e: print(f, "fail", e)
PY
