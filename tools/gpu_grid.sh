#!/bin/bash
# cell-grid experiments on C4/C3: parity tests, then bench lines over grid densities
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cell_grid.py -x -q > gpurun_out/grid_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/grid_pytest.log
tail -15 gpurun_out/grid_pytest.log
B="python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu-baseline"
timeout 300 $B --accel bvh > gpurun_out/g_bvh.json 2> gpurun_out/g_bvh.err
for d in 0.25 0.5 1 2 4; do
  ERT_CELL_GRID_DENSITY=$d timeout 300 $B --accel grid > gpurun_out/g_d$d.json 2> gpurun_out/g_d$d.err
done
ERT_CELLS_FROM=1 timeout 300 $B --accel grid > gpurun_out/g_from1.json 2> gpurun_out/g_from1.err
ERT_CELLS_FROM=2 timeout 300 $B --accel grid > gpurun_out/g_from2.json 2> gpurun_out/g_from2.err
ERT_CELLS_REFILL_FROM=1 timeout 300 $B --accel grid > gpurun_out/g_refill1.json 2> gpurun_out/g_refill1.err
ERT_CELLS_REFILL_FROM=9 timeout 300 $B --accel grid > gpurun_out/g_refill9.json 2> gpurun_out/g_refill9.err
timeout 300 python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline --accel grid > gpurun_out/g_c3.json 2> gpurun_out/g_c3.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/g_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; fr=r["frame"]
        print(f.split('/')[-1], "accel", d["config"]["accel"], "ms %.3f"%d["ms_per_step"], "path %.2f shadow %.2f other %.2f"%(fr["ms"]["path"],fr["ms"]["shadow"],fr["ms"]["other"]), "cells", r.get("cell_steps"), "filt", r["sphere_filter_tests"], "box", r["box_tests"], "exact", r["exact_fp64_sphere_tests"], "upload %.2f"%d["config"]["scene_upload_s"])
    except Exception as e: print(f, "fail", e)
PY
