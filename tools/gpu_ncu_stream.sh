#!/bin/bash
# kernel table (launch list, gpu__time_duration) + ncu --set full of the streaming kernels at 512^3
mkdir -p gpurun_out
python tools/kernel_table.py run 512 > gpurun_out/kt_plain.log 2>&1 || { tail -5 gpurun_out/kt_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_kernel_table_launches.csv python tools/kernel_table.py run 512 > gpurun_out/kt_ncu.log 2>&1
python tools/kernel_table.py report gpurun_out/r2_kernel_table_launches.csv 512 > gpurun_out/r2_kernel_table.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fetch_stats|k_histogram|k_bilateral|k_sdf_unbrick|k_boxavg|k_lin_field|k_sdf_assemble|k_sdf_events|k_sdf_band|k_clip" -c 14 -o gpurun_out/r2_stream -f python tools/kernel_table.py run 512 > gpurun_out/kt_ncu2.log 2>&1
ncu -i gpurun_out/r2_stream.ncu-rep --page raw --csv > gpurun_out/r2_stream_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_stream_raw.csv > gpurun_out/r2_stream_summary.txt 2>&1
cat gpurun_out/r2_kernel_table.txt
