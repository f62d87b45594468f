#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/s_probe2.txt
python tools/sdf_probe.py 512 >> gpurun_out/s_probe2.txt 2>&1
export VR_LIB=tools/ab/libvr_ab.so
VR_SDF_WAVE=6 python tools/sdf_probe.py 512 >> gpurun_out/s_probe2.txt 2>&1
for v in 0 1 2 3; do
  VR_SDF_W6=$v python tools/sdf_probe.py 512 >> gpurun_out/s_probe2.txt 2>&1
  VR_SDF_W6=$v python tools/sdf_probe.py 1000,200,136 >> gpurun_out/s_probe2.txt 2>&1
done
cat gpurun_out/s_probe2.txt
