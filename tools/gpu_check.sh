#!/bin/bash
# quick correctness (wavefront / grid / parity tests) + frame times of C4 and C3 + one part of 8
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_cell_grid.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_erl_reference.py -x -q > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_pytest.log
tail -4 gpurun_out/check_pytest.log
python tools/frame_once.py c4 5 2>&1 | tail -2
python tools/frame_once.py c3 4 2>&1 | tail -1
python tools/part_probe.py c4 1 10 2>&1 | head -1
python tools/part_probe.py c3 1 10 2>&1 | head -1
python tools/part_probe.py c4 8 10 2>&1 | head -1
python tools/part_probe.py c4 4 10 2>&1 | head -1
python tools/part_probe.py c4 2 10 2>&1 | head -1
