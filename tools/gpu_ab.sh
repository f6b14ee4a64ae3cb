#!/bin/bash
# A/B of library build variants: cell-grid parity tests, then bench lines (WORKLOAD, default c4) over
# "variant:density[:ENV=VAL]" combos; variant = base or the suffix of eraytracer_b200/lib/libert_b200_<variant>.so
mkdir -p gpurun_out; rm -f gpurun_out/h_*.json gpurun_out/h_*.err
timeout 900 python -m pytest tests/test_gpu_cell_grid.py -x -q > gpurun_out/grid_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/grid_pytest.log
tail -5 gpurun_out/grid_pytest.log
W=${WORKLOAD:-c4}
B="python bench.py --workload $W --steps 5 --warmup 2 --no-cpu-baseline --accel grid"
for combo in "$@"; do
  IFS=: read v d e <<< "$combo"
  lib=$PWD/eraytracer_b200/lib/libert_b200.so; [ "$v" != base ] && lib=$PWD/eraytracer_b200/lib/libert_b200_$v.so
  tag=${v}_d${d}_${e//=/}
  env ERT_B200_LIB=$lib ERT_CELL_GRID_DENSITY=$d ${e:-X=1} timeout 300 $B > gpurun_out/h_$tag.json 2> gpurun_out/h_$tag.err || tail -3 gpurun_out/h_$tag.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/h_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; fr=r["frame"]
        print(f.split('/')[-1], "accel", d["config"]["accel"], "ms %.3f"%d["ms_per_step"], "path %.2f shadow %.2f other %.2f"%(fr["ms"]["path"],fr["ms"]["shadow"],fr["ms"]["other"]), "cells", r.get("cell_steps"), "filt", r["sphere_filter_tests"], "exact", r["exact_fp64_sphere_tests"])
    except Exception as e: print(f, "fail", e)
PY
