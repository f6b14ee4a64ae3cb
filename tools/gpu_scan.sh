#!/bin/bash
# brute-force scan kernel: parity tests that use accel=linear, then bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cell_grid.py tests/test_gpu_edges.py -x -q -k "linear or c3 or c4 or golden or synthetic or accel" > gpurun_out/scan_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/scan_pytest.log
tail -5 gpurun_out/scan_pytest.log
timeout 600 python bench.py --workload c3 --accel linear --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scan_c3_linear.json 2> gpurun_out/scan_c3_linear.err; echo "c3 linear rc=$?"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scan_c4.json 2> gpurun_out/scan_c4.err; echo "c4 rc=$?"
python - <<PY
import json
for f in ("scan_c3_linear","scan_c4"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); r=d["roofline"]; print(f, "ms", d["ms_per_step"], "frac", r["frac"], r["frame"]["class_frac"], r["frame"]["ms"]); print(json.dumps(d.get("roofline_scan"),indent=1))
    except Exception as e: print(f, "fail", e)
PY
