#!/bin/bash
# N-GPU run: D2H ceiling probe, bench lines of C4 and C5 (N = $1, default 8)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29611 tools/d2h_probe.py > gpurun_out/d2h_probe_n$N.txt 2> gpurun_out/d2h_probe_n$N.err; echo "probe rc=$?"; cat gpurun_out/d2h_probe_n$N.txt
$TR --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 --workload c4 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "c4 rc=$?"
$TR --master-port 29631 bench.py --gpus $N --steps 64 --warmup 8 --workload c5 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "c5 rc=$?"
python - <<PY
import json
for f in ("bench_c4_n$N","bench_c5_n$N"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); e=d["e2e"]
        print(f, "n", d["n_gpus"], "ms %.3f Mrays/s %.0f e2e %.0f fps %.1f d2h/gpu %.1f GB/s (%.2f of %.1f) assembled %s"%(d["ms_per_step"], d["value"], e["value"], e["frames_per_s"], e["d2h_gbs_per_gpu"], e["roofline"]["frac"], e["roofline"]["peak"], d["config"].get("assembled_frame_equals_single_gpu")))
    except Exception as ex: print(f, "fail", ex)
PY
tail -3 gpurun_out/bench_c4_n$N.err
