#!/bin/bash
# whole single-GPU suite, host-side breakdown of the headless job, default bench line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log | cut -c1-300
python tools/e2e_probe.py > gpurun_out/f_e2e_probe.txt 2>&1; tail -5 gpurun_out/f_e2e_probe.txt
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench exit $?"; tail -3 gpurun_out/f_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/f_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'sdf', d.get('sdf_build_ms'), 'lin', d['hw_linear']['value'], d['hw_linear']['e2e']['value'])
        print('closeup', d['closeup']['value'], 'per_frame', d['per_frame_schedule']['value'], 'interactive', d['interactive_loop']['value'])
PY
