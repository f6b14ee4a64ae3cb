#!/bin/bash
# bench lines of every workload (N=1) + the reference arm
mkdir -p gpurun_out
for w in c4 c4d1 c3 c2 c2d5 c5 c1; do
  timeout 600 python bench.py --workload $w --steps ${STEPS:-10} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || { echo "bench $w failed"; tail -5 gpurun_out/bench_$w.err; }
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c*.json")):
    try:
        d=json.load(open(f)); r=d.get("roofline",{}); e=d["e2e"]; s=d.get("roofline_scan") or {}
        print(f.split('/')[-1], "ms %.3f Mrays/s %.0f e2e %.0f (%.3f ms, d2h %.1f GB/s = %.2f of %.1f) roofline %s frac %.3f scan %s cpu %s"%(d["ms_per_step"], d["value"], e["value"], e["ms_per_step"], e["d2h_gbs_per_gpu"], e["roofline"]["frac"], e["roofline"]["peak"], r.get("bound"), r.get("frac"), s.get("frac"), d.get("cpu_baseline",{}).get("value")))
    except Exception as e: print(f, "fail", e)
PY
