#!/bin/bash
# SDF build: parity tests, then the level kernels / tile geometries side by side (A/B build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharding.py -m gpu -q -k "sdf or slab" --timeout=600 -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -5 gpurun_out/s_pytest.log | cut -c1-200
: > gpurun_out/s_probe.txt
python tools/sdf_probe.py 512 >> gpurun_out/s_probe.txt 2>&1
python tools/sdf_probe.py 1024,200,136 >> gpurun_out/s_probe.txt 2>&1
python tools/sdf_probe.py 256 >> gpurun_out/s_probe.txt 2>&1
export VR_LIB=tools/ab/libvr_ab.so
VR_SDF_PDL=0 python tools/sdf_probe.py 512 >> gpurun_out/s_probe.txt 2>&1
for v in 1 2 3 4 5; do
  VR_SDF_VARIANT=$v python tools/sdf_probe.py 512 >> gpurun_out/s_probe.txt 2>&1
done
cat gpurun_out/s_probe.txt
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none -k regex:"k_sdf_count|k_sdf_assemble" -c 4 --csv --log-file gpurun_out/s_w9_launches.csv python tools/sdf_probe.py 512 > gpurun_out/s_ncu1.log 2>&1
grep -v "^==" gpurun_out/s_w9_launches.csv | cut -d, -f5,13,15 | head -30
