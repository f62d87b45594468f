#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "linear or quiet or hw_linear" --timeout=600 -p no:cacheprovider 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_lin_" -c 8 --csv --log-file gpurun_out/l_launches.csv python tools/kernel_table.py run 512 > gpurun_out/l_kt.log 2>&1
grep "k_lin" gpurun_out/l_launches.csv | cut -d, -f5,15 | cut -c1-120
timeout 300 python tools/lin_probe.py 512 quick 2>/dev/null | cut -c1-300 | head -4
