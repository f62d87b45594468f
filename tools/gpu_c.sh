#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
timeout 300 python tools/lin_probe.py 512 quick > gpurun_out/c_lin_probe.jsonl 2> gpurun_out/c_lin_probe.err
timeout 300 python tools/interactive_probe.py 512 > gpurun_out/c_interactive.jsonl 2> gpurun_out/c_interactive.err
bash tools/gpu_ncu_stream.sh > gpurun_out/c_stream.log 2>&1
tail -3 gpurun_out/c_pytest.log; cat gpurun_out/c_lin_probe.jsonl | cut -c1-250; cat gpurun_out/c_interactive.jsonl
