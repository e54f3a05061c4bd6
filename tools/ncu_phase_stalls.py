"""Stall reasons per source-line range from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_phase_stalls.py source.csv main_file_prefix name:lo:hi ..."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
prefix = sys.argv[2]
phases = {}
for a in sys.argv[3:]:
    k, lo, hi = a.split(":"); phases[k] = (int(lo), int(hi))
hdr = next(r for r in rows if len(r) > 10 and r[0] == 'Line No')
st = [(n, i) for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n and 'not_issued' not in n]
cur = None
agg = defaultdict(lambda: defaultdict(int))
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        l = int(r[0]); key = 'other:' + cur
        if cur.startswith(prefix):
            for k, (a, b) in phases.items():
                if a <= l <= b:
                    key = k; break
        for n, i in st:
            try: agg[key][n] += int(r[i])
            except (ValueError, IndexError): pass
tot = sum(sum(v.values()) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1].values())):
    s = sum(v.values())
    if s == 0: continue
    top = sorted(v.items(), key=lambda kv: -kv[1])[:4]
    print(f"{k:30s} {s / tot:6.1%}  " + ", ".join(f"{n[6:]} {c / s:.0%}" for n, c in top))
