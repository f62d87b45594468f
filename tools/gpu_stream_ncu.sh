#!/bin/bash
# hw-linear / volume-op tests, then full captures of the streaming kernels at 512^3 (one launch each)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_ref_opencl_gpu.py -m gpu -q -x -k "linear or sampling or stats or hist or bilateral or quiet" --timeout=600 -p no:cacheprovider 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdf_count|k_sdf_assemble8|k_sdf_events|k_histogram_lut|k_fetch_stats_v8i|k_lin_corners|k_lin_cells|k_bilateral|k_boxavg|k_clip|k_sdf_unbrick" -c 14 -o gpurun_out/z_stream2 -f python tools/kernel_table.py run 512 > gpurun_out/z_ncu4.log 2>&1
ncu -i gpurun_out/z_stream2.ncu-rep --page raw --csv > gpurun_out/z_stream2_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/z_stream2_raw.csv > gpurun_out/z_stream2_summary.txt 2>&1
rm -f gpurun_out/z_stream2.ncu-rep
grep "^---\|time_duration\|issue_active\|dram_throughput" gpurun_out/z_stream2_summary.txt
