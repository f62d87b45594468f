#!/bin/bash
mkdir -p gpurun_out
python tools/profile_target.py 64 512 default nearest 2 > gpurun_out/ncu_near_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_pt|k_primary" -c 2 -o gpurun_out/r2_near_trace_64frames -f python tools/profile_target.py 64 512 default nearest 2 > gpurun_out/ncu_near.log 2>&1
ncu -i gpurun_out/r2_near_trace_64frames.ncu-rep --page raw --csv > gpurun_out/r2_near_trace_64frames_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_near_trace_64frames_raw.csv > gpurun_out/r2_near_trace_64frames_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_pt" -c 1 -o gpurun_out/r2_lin_trace_v2 -f python tools/profile_target.py 64 512 default linear 2 > gpurun_out/ncu_lin_v2.log 2>&1
ncu -i gpurun_out/r2_lin_trace_v2.ncu-rep --page raw --csv > gpurun_out/r2_lin_trace_v2_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_lin_trace_v2_raw.csv > gpurun_out/r2_lin_trace_v2_summary.txt 2>&1
cat gpurun_out/r2_near_trace_64frames_summary.txt | head -60
