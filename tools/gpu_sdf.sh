#!/bin/bash
# SDF build: parity tests, then the level kernels / tile geometries side by side (A/B build)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sharding.py -m gpu -q -x -k "sdf or slab" --timeout=300 -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -5 gpurun_out/s_pytest.log | cut -c1-200
: > gpurun_out/s_probe.txt
timeout 120 python tools/sdf_probe.py 512 >> gpurun_out/s_probe.txt 2>&1
timeout 120 python tools/sdf_probe.py 1024,200,136 >> gpurun_out/s_probe.txt 2>&1
timeout 120 python tools/sdf_probe.py 256 >> gpurun_out/s_probe.txt 2>&1
timeout 120 python tools/sdf_probe.py 1024 >> gpurun_out/s_probe.txt 2>&1
export VR_LIB=tools/ab/libvr_ab.so
VR_SDF_FLOW=0 timeout 120 python tools/sdf_probe.py 512 >> gpurun_out/s_probe.txt 2>&1
VR_SDF_FLOW=0 timeout 120 python tools/sdf_probe.py 1024 >> gpurun_out/s_probe.txt 2>&1
for v in 1 2 3 4 5; do
  VR_SDF_VARIANT=$v timeout 120 python tools/sdf_probe.py 512 >> gpurun_out/s_probe.txt 2>&1
done
cat gpurun_out/s_probe.txt
