#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "hist or render_tf or stats" --timeout=600 -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest.log
tail -12 gpurun_out/h_pytest.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_hist|k_fetch_stats|k_bilateral|k_boxavg|k_lin_field|k_clip|k_tf_" -c 60 --csv --log-file gpurun_out/h_launches.csv python tools/kernel_table.py run 512 > gpurun_out/h_kt.log 2>&1
python tools/kernel_table.py report gpurun_out/h_launches.csv 512
