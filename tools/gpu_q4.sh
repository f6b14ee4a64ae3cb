#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/q4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q4_pytest.log; tail -4 gpurun_out/q4_pytest.log
for e in ERT_NO_Q4=1 ERT_Q4_FROM=1 ERT_Q4_FROM=0 ERT_Q4_FROM=2; do
  env $e timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/q4_$e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$e ms %.2f upload %.2f path_ms %.2f box %.3g filt %.3g'%(d['ms_per_step'], d['config']['scene_upload_s'], r['frame']['ms']['path'], r['frame']['box_tests'], r['frame']['sphere_filter_tests']))" || tail -3 gpurun_out/q4_$e.err
done
ERT_Q4_FROM=1 timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 q4 ms %.2f'%d['ms_per_step'])"
