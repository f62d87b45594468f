#!/bin/bash
# one full capture of two mid-build levels of the default SDF level kernel at 512^3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_sdf_wave8" -s 30 -c 1 -o gpurun_out/s_w8_l31 -f python tools/sdf_probe.py 512 > gpurun_out/s_ncu3.log 2>&1
ncu -i gpurun_out/s_w8_l31.ncu-rep --page raw --csv > gpurun_out/s_w8_l31_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/s_w8_l31_raw.csv > gpurun_out/s_w8_l31_summary.txt 2>&1
ncu -i gpurun_out/s_w8_l31.ncu-rep --page source --csv > gpurun_out/s_w8_l31_source.csv 2>/dev/null
cat gpurun_out/s_w8_l31_summary.txt
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none -k regex:"k_sdf_wave8" -c 125 --csv --log-file gpurun_out/s_w8_launches.csv python tools/sdf_probe.py 512 > gpurun_out/s_ncu1.log 2>&1
