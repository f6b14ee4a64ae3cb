#!/bin/bash
# scan-kernel probe (FFMA2 rate, filter-loop variants) + a baseline bench line
mkdir -p gpurun_out
timeout 300 tools/_build/scan_probe > gpurun_out/scan_probe.txt 2>&1; echo "probe rc=$?" >> gpurun_out/scan_probe.txt
cat gpurun_out/scan_probe.txt
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_c4.json 2> gpurun_out/r02_base_c4.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_base_c4.json")); print(d["ms_per_step"], d["roofline"]["frame"]["ms"], d["e2e"]["ms_per_step"])
PY
