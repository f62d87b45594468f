#!/bin/bash
# round-2 final single-GPU pass: suite, bench (both arms), launch lists, full captures of the top kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/z_gpus.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/z_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/z_pytest.log
tail -4 gpurun_out/z_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_bench_reference_arm.json 2>> gpurun_out/z_bench.err; echo "ref exit $?"
# launch list of the bench command (step kernels), cold-cache and serialised: shares, not absolutes
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_primary|k_trace_pt|k_resolve|k_cache_reset" -c 400 --csv --log-file gpurun_out/z_launches_bench_step_kernels.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-big > gpurun_out/z_ncu_bench.log 2>&1
# every kernel family once at 512^3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/z_kernel_table_launches.csv python tools/kernel_table.py run 512 > gpurun_out/z_kt.log 2>&1
python tools/kernel_table.py report gpurun_out/z_kernel_table_launches.csv 512 > gpurun_out/z_kernel_table.txt 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace_pt|k_primary" -c 2 -o gpurun_out/z_near_trace -f python tools/profile_target.py 64 512 default nearest 2 > gpurun_out/z_ncu1.log 2>&1
ncu -i gpurun_out/z_near_trace.ncu-rep --page raw --csv > gpurun_out/z_near_trace_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/z_near_trace_raw.csv > gpurun_out/z_near_trace_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace_pt|k_primary" -c 2 -o gpurun_out/z_lin_trace -f python tools/profile_target.py 64 512 default linear 2 > gpurun_out/z_ncu2.log 2>&1
ncu -i gpurun_out/z_lin_trace.ncu-rep --page raw --csv > gpurun_out/z_lin_trace_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/z_lin_trace_raw.csv > gpurun_out/z_lin_trace_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdf_wave9|k_sdf_count|k_sdf_assemble8|k_sdf_events|k_histogram_lut|k_fetch_stats_v8i|k_lin_corners|k_lin_cells|k_bilateral" -s 40 -c 96 -o gpurun_out/z_stream -f python tools/kernel_table.py run 512 > gpurun_out/z_ncu3.log 2>&1
ncu -i gpurun_out/z_stream.ncu-rep --page raw --csv > gpurun_out/z_stream_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/z_stream_raw.csv > gpurun_out/z_stream_summary_all.txt 2>&1
rm -f gpurun_out/z_stream.ncu-rep
python - <<'PY'
import json
for l in open('gpurun_out/z_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'sdf', d.get('sdf_build_ms'), 'lin', d['hw_linear']['value'], d['hw_linear']['e2e']['value'])
        print('closeup', d['closeup']['value'], 'per_frame', d['per_frame_schedule']['value'], 'interactive', d['interactive_loop']['value'], 'ref_on_gpu', (d.get('reference_on_gpu') or {}).get('value'))
PY
cat gpurun_out/z_kernel_table.txt
