#!/bin/bash
# Round evidence in one call: bench line, ncu launch list of the same command, --set full captures of the
# path kernels, the shadow+shade kernel and the brute-force scan (each after its plain run exited 0).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
echo "launch list rc=$?"
python tools/frame_once.py c4 3 > gpurun_out/frame_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wf_trace_path" -c 5 -f -o gpurun_out/prof_path python tools/frame_once.py c4 1 > gpurun_out/frame_ncu_path.log 2>&1
echo "path rc=$?"
python tools/ncu_summary.py gpurun_out/prof_path.ncu-rep gpurun_out/r02_c4_grid_path_full.txt "one C4 frame (3840x2160, depth 5, 1M spheres): the five cell-grid path launches" > /dev/null
python tools/ncu_lines.py gpurun_out/prof_path.ncu-rep wf_trace_path_refill 400 > gpurun_out/r02_c4_refill_lines.txt 2>&1
rm -f gpurun_out/prof_path.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"wf_shadow_shade|wf_finalize" -c 6 -f -o gpurun_out/prof_ss python tools/frame_once.py c4 1 > gpurun_out/frame_ncu_ss.log 2>&1
echo "shadow+shade rc=$?"
python tools/ncu_summary.py gpurun_out/prof_ss.ncu-rep gpurun_out/r02_c4_shadow_shade_full.txt "one C4 frame: the five wf_shadow_shade launches and wf_finalize" > /dev/null
python tools/ncu_lines.py gpurun_out/prof_ss.ncu-rep wf_shadow_shade 400 > gpurun_out/r02_c4_shadow_shade_lines.txt 2>&1
rm -f gpurun_out/prof_ss.ncu-rep
python tools/scan_band.py c4 2 > gpurun_out/scan_band_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wf_scan -c 4 -f -o gpurun_out/prof_scan_c4 python tools/scan_band.py c4 1 > gpurun_out/scan_ncu_full.log 2>&1
echo "scan rc=$?"
python tools/ncu_summary.py gpurun_out/prof_scan_c4.ncu-rep gpurun_out/r02_c4_scan_full.txt "rows 1080-1083 of the C4 frame, accel=linear: wf_scan launches (path bounce 0, shadow, path bounce 1, shadow)" > /dev/null
python tools/ncu_lines.py gpurun_out/prof_scan_c4.ncu-rep wf_scan 80 > gpurun_out/r02_c4_scan_lines.txt 2>&1
rm -f gpurun_out/prof_scan_c4.ncu-rep
tail -2 gpurun_out/frame_plain.log; tail -2 gpurun_out/scan_band_plain.log
