#!/bin/bash
# full capture of the dataflow level kernel at 512^3
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdf_flow" -c 1 -o gpurun_out/s_flow -f python tools/sdf_probe.py 512 > gpurun_out/s_ncu3.log 2>&1
ncu -i gpurun_out/s_flow.ncu-rep --page raw --csv > gpurun_out/s_flow_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/s_flow_raw.csv > gpurun_out/s_flow_summary.txt 2>&1
ncu -i gpurun_out/s_flow.ncu-rep --page source --csv > gpurun_out/s_flow_source.csv 2>/dev/null
cat gpurun_out/s_flow_summary.txt
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/s_flow_raw.csv')))
h,u,r=rows[0],rows[1],rows[2]
for i,n in enumerate(h):
    if 'pcsamp_warps_issue_stalled' in n and 'not_issued' not in n and r[i] not in ('','0'):
        print(n.split('stalled_')[1], r[i])
for n in ['smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.per_cycle_active']:
    if n in h: print(n, r[h.index(n)])
PY
