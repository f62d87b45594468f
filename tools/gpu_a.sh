#!/bin/bash
# round-2 GPU call A: full GPU test suite, hw-linear A/B probe, bench (no big legs)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
timeout 600 python tools/lin_probe.py 512 > gpurun_out/a_lin_probe.jsonl 2> gpurun_out/a_lin_probe.err
VR_LIB=$PWD/tools/ab/libvr_ab.so timeout 300 python tools/lin_probe.py 512 quick > gpurun_out/a_lin_probe_ab.jsonl 2>> gpurun_out/a_lin_probe.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?" >> gpurun_out/a_bench.err
tail -5 gpurun_out/a_pytest.log
