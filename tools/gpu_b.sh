#!/bin/bash
# round-2 GPU call B: GPU tests, A/B probes (product + A/B library), bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
VR_LIB=$PWD/tools/ab/libvr_ab.so timeout 600 python tools/lin_probe.py 512 > gpurun_out/b_lin_probe.jsonl 2> gpurun_out/b_lin_probe.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench exit $?" >> gpurun_out/b_bench.err
tail -4 gpurun_out/b_pytest.log
