#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharding.py tests/test_ref_opencl_gpu.py -m gpu -q -x -k "bilateral or filter or slab or volume" --timeout=600 -p no:cacheprovider 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_bilateral" -c 4 --csv --log-file gpurun_out/b_launches.csv python tools/kernel_table.py run 512 > gpurun_out/b_kt.log 2>&1
grep "k_bilateral" gpurun_out/b_launches.csv | awk -F'","' '{print substr($5,1,16), $(NF)}'
