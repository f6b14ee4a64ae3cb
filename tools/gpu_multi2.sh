#!/bin/bash
# N-GPU bench lines: C4 always, C5 when $2 = c5 (N = $1)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "c4 rc=$?"
if [ "$2" = c5 ]; then
$TR --master-port 29631 bench.py --gpus $N --steps 64 --warmup 8 --workload c5 --no-cpu-baseline > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "c5 rc=$?"
fi
python - <<PY
import json
for f in ("bench_c4_n$N","bench_c5_n$N"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); e=d["e2e"]
        print(f, "n", d["n_gpus"], "ms %.3f Mrays/s %.0f e2e %.0f fps %.1f d2h/gpu %.1f GB/s assembled %s"%(d["ms_per_step"], d["value"], e["value"], e["frames_per_s"], e["d2h_gbs_per_gpu"], d["config"].get("assembled_frame_equals_single_gpu")))
    except Exception as ex: print(f, "fail", ex)
PY
tail -2 gpurun_out/bench_c4_n$N.err
