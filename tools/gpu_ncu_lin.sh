#!/bin/bash
# ncu --set full of the hw-linear trace phase (k_primary + k_trace_pt, 64 frames) and of the per-frame schedule (k_trace + k_trace_pt, 4 frames)
mkdir -p gpurun_out
python tools/profile_target.py 64 512 default linear 2 > gpurun_out/ncu_lin_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_pt|k_primary|k_lin_field" -c 3 -o gpurun_out/r2_lin_trace_64frames -f python tools/profile_target.py 64 512 default linear 2 > gpurun_out/ncu_lin.log 2>&1
ncu -i gpurun_out/r2_lin_trace_64frames.ncu-rep --page raw --csv > gpurun_out/r2_lin_trace_64frames_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_lin_trace_64frames_raw.csv > gpurun_out/r2_lin_trace_64frames_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace" -c 2 -o gpurun_out/r2_lin_perframe -f python tools/profile_target.py 1 512 default linear 1 > gpurun_out/ncu_lin1.log 2>&1
ncu -i gpurun_out/r2_lin_perframe.ncu-rep --page raw --csv > gpurun_out/r2_lin_perframe_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_lin_perframe_raw.csv > gpurun_out/r2_lin_perframe_summary.txt 2>&1
ncu -i gpurun_out/r2_lin_trace_64frames.ncu-rep --page source --csv --kernel-name regex:k_trace_pt > gpurun_out/r2_lin_trace_pt_source.csv 2>/dev/null
ls -la gpurun_out | tail -12
