"""Per-source-line instruction / stall totals from `ncu --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py export.csv [top]"""
import csv, sys, re, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
fname = ''
out = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = os.path.basename(r[1]); continue
    if r[0] == 'Line No':
        hdr = r
        ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed')
        ss = hdr.index('Warp Stall Sampling (All Samples)')
        continue
    if hdr is None or not re.fullmatch(r'\d+', r[0]) or len(r) <= te:
        continue
    try:
        out.append((fname, int(r[0]), r[1].strip(), int(r[ie]), int(r[te]), int(r[ss])))
    except ValueError:
        pass
tot = [sum(o[i] for o in out) for i in (3, 4, 5)]
print(f"total warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  avg threads/inst {tot[1]/max(tot[0],1):.1f}  stall samples {tot[2]:,}")
for f, ln, src, a, b, c in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{f[:14]:14s}:{ln:4d} inst {100*a/max(tot[0],1):5.1f}%  thr/inst {b/max(a,1):5.1f}  stall {100*c/max(tot[2],1):5.1f}%  | {src[:105]}")
