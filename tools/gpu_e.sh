#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
timeout 400 python tools/lin_probe.py 512 quick > gpurun_out/e_lin_probe.jsonl 2> gpurun_out/e_lin_probe.err
tail -15 gpurun_out/e_pytest.log | cut -c1-200; python - <<'PY'
import json
for l in open('gpurun_out/e_lin_probe.jsonl'):
    r=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in r.items() if 'flush' not in k and 'gsamples' not in k})
PY
