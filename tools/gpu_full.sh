#!/bin/bash
# full GPU test suite + smoke + bench lines of every workload (N=1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -4 gpurun_out/full_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/full_smoke.log; tail -5 gpurun_out/full_smoke.log
for w in c4 c4d1 c3 c2 c2d5 c5; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || { echo "bench $w failed"; tail -5 gpurun_out/bench_$w.err; }
done
timeout 600 python bench.py --workload c4 --accel bvh > gpurun_out/bench_c4_bvh.json 2> gpurun_out/bench_c4_bvh.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c*.json"))+["gpurun_out/bench_ref_c4.json"]:
    try:
        d=json.load(open(f)); r=d.get("roofline",{}); print(f.split('/')[-1], "ms %.3f Mrays/s %.0f e2e %.0f frac %s cpu %s clocks %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("frac"), d.get("cpu_baseline",{}).get("value"), d.get("clocks")))
    except Exception as e: print(f, "fail", e)
PY
