#!/bin/bash
# per-launch device times (ncu, cold cache, serialised) of one C4 frame: whole frame and one part of 8
mkdir -p gpurun_out
python tools/part_probe.py c4 ${1:-8} 2 > gpurun_out/part_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 21 --csv --log-file gpurun_out/part_launches.csv python tools/part_probe.py c4 ${1:-8} 2 > gpurun_out/part_ncu.log 2>&1
tail -2 gpurun_out/part_plain.log
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/part_launches.csv")) if len(r)>5 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4].split("(")[0][-60:]; v=float(r[-1].replace(",","")); unit=r[-2]
    if unit=="ns": v/=1e6
    elif unit in ("us","usecond"): v/=1e3
    tot+=v
    print("%-62s %8.4f ms"%(name,v))
print("total %.3f ms"%tot)
PY
