#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/s_probe2.txt
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sharding.py tests/test_big_gpu.py -m gpu -q -x -k "sdf or slab or 1024" --timeout=300 -p no:cacheprovider 2>&1 | tail -3
for n in 256 512 640 1024; do timeout 120 python tools/sdf_probe.py $n >> gpurun_out/s_probe2.txt 2>&1; done
cat gpurun_out/s_probe2.txt
