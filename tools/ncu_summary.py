"""Summarise an `ncu --page raw --csv` export: python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_atom.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
for r in rows[2:]:
    print('---', r[h.index('Kernel Name')][:60], r[h.index('Grid Size')] if 'Grid Size' in h else '')
    for w in want:
        if w in h:
            print(f"  {w:72s} {r[h.index(w)]:>18s} {units[h.index(w)]}")
    st = [(float(r[i].replace(',', '')), n) for i, n in enumerate(h)
          if 'issue_stalled' in n and n.endswith('_per_warp_active.pct') and r[i]]
    for v, n in sorted(st, reverse=True)[:7]:
        print(f"  stall {n.split('issue_stalled_')[1].split('_per_warp')[0]:40s} {v:8.2f} %")
