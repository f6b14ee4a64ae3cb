#!/bin/bash
# first GPU pass of the wavefront path: parity tests, then C4/C3 timings with BVH builder knobs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
for cfg in "4 0" "8 1" "8 2" "8 4" "6 2"; do
  set -- $cfg
  ERT_BVH_LEAF_MAX=$1 ERT_BVH_TRAV_COST=$2 timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_c4_leaf$1_ct$2.json 2> gpurun_out/s1_c4_leaf$1_ct$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s1_c4_leaf$1_ct$2.json"))
    print("c4 leaf$1 ct$2", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["box_tests"], d["roofline"]["sphere_filter_tests"])
except Exception as e: print("fail", e)
PY
done
timeout 300 python bench.py --workload c4 --accel bvh_mega --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1_c4_mega.json 2> gpurun_out/s1_c4_mega.err
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_c3.json 2> gpurun_out/s1_c3.err
ERT_BVH_LEAF_MAX=8 ERT_BVH_TRAV_COST=2 timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_c3_leaf8_ct2.json 2> gpurun_out/s1_c3_leaf8.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s1_c*.json")):
    try:
        d=json.load(open(f)); print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], d["gpu_launches"])
    except Exception as e: print(f, "fail", e)
PY
