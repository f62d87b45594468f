#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
timeout 300 python tools/lin_probe.py 512 quick > gpurun_out/d_lin_probe.jsonl 2> gpurun_out/d_lin_probe.err
timeout 300 python tools/interactive_probe.py 512 > gpurun_out/d_interactive.jsonl 2> gpurun_out/d_interactive.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench exit $?" >> gpurun_out/d_bench.err
tail -3 gpurun_out/d_pytest.log; cat gpurun_out/d_lin_probe.jsonl | cut -c1-250; cat gpurun_out/d_interactive.jsonl | cut -c1-330; tail -2 gpurun_out/d_bench.err
