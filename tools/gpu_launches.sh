#!/bin/bash
# per-launch kernel times of one steady-state frame: tools/gpu_launches.sh <workload> <accel> <tag>
mkdir -p gpurun_out
CMD="python bench.py --workload $1 --accel $2 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 44 --csv --log-file gpurun_out/launches_$3.csv $CMD > gpurun_out/launches_$3.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$3.csv')) if len(r)>10 and r[0].isdigit()]
names=[(r[4].split('(')[0][-40:], float(r[-1])/1e6) for r in rows]
start=[i for i,(n,t) in enumerate(names) if 'wf_trace_path<1' in n or 'wf_scan_path<1' in n]
i0=start[0]; i1=start[1] if len(start)>1 else len(names)
for n,t in names[i0:i1]: print("%-45s %.3f"%(n,t))
PY
