"""profiles/traffic_*.json and profiles/issue_*.json (read by bench.py for roofline.traffic / roofline.issue_bound) from the raw export
of an `ncu --set full` capture of one 64-frame trace phase:  python tools/ncu_to_bench_profiles.py raw.csv nearest|linear source-note"""
import csv, json, sys, os
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
which = sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
def val(r, name):
    i = h.index(name)
    v = float(r[i].replace(',', ''))
    u = units[i].lower()
    scale = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}.get(u, 1)
    return v * scale
kern, total = {}, 0.0
issue = None
for r in rows[2:]:
    name = r[h.index('Kernel Name')].replace('void ', '').split('(')[0]
    rd, wr, ms = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum'), val(r, 'gpu__time_duration.sum')
    kern[name] = {"dram_read": int(rd), "dram_write": int(wr), "ms": round(ms, 4)}
    total += rd + wr
    if name.startswith('k_trace_pt'):
        ia = val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active') / 100.0
        lanes = val(r, 'smsp__thread_inst_executed_per_inst_executed.ratio')
        issue = {"kernel": name, "issue_active": round(ia, 4), "lanes_per_instruction": round(lanes, 2), "frac_of_lane_issue_peak": round(ia * lanes / 32.0, 4),
                 "warp_instructions": int(val(r, 'smsp__inst_executed.sum')), "l1tex_throughput_pct": val(r, 'l1tex__throughput.avg.pct_of_peak_sustained_active'),
                 "lts_throughput_pct": val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
                 "what": "second bound beside the notional HBM fraction: issue-active x active lanes per instruction / 32 of k_trace_pt", "source": note}
suffix = "" if which == "nearest" else "_lin"
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
json.dump({"dram_bytes_per_launch": int(total), "kernels": kern,
           "launch": "one trace phase of 64 frames at 1920x1080 on the 512^3 bench scene (vr_render_frames, default camera, trace mode 2), sampling " + which,
           "source": note}, open(os.path.join(root, f"traffic_k_trace{suffix}.json"), "w"), indent=1)
json.dump(issue, open(os.path.join(root, f"issue_k_trace_pt{suffix}.json"), "w"), indent=1)
print(json.dumps(issue)); print(int(total))
