"""Attribute warp instructions and stall samples of an ncu source page to line ranges.
usage: ncu_phases.py source.csv rows file_prefix name:lo:hi ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nrows = int(sys.argv[2]); prefix = sys.argv[3]
ph = {}
for a in sys.argv[4:]:
    k, lo, hi = a.split(":"); ph[k] = (int(lo), int(hi))
cur = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit() and r[7].isdigit():
        data.append((int(r[7]), int(r[6]) if r[6].isdigit() else 0, cur, int(r[0]), r[1].strip()[:90]))
tots = sum(d[1] for d in data); toti = sum(d[0] for d in data)
acc = {k: [0, 0] for k in ph}; other = [0, 0]
for n, s, f, l, _ in data:
    hit = False
    if f.startswith(prefix):
        for k, (a, b) in ph.items():
            if a <= l <= b: acc[k][0] += n; acc[k][1] += s; hit = True; break
    if not hit: other[0] += n; other[1] += s
print(f"total warp-instr/row {toti/nrows:.0f}, samples {tots}")
for k, v in acc.items(): print(f"{k:10s} instr {v[0]/toti:6.1%} ({v[0]/nrows:7.0f}/row) samples {v[1]/tots:6.1%}")
print(f"{'other':10s} instr {other[0]/toti:6.1%} samples {other[1]/tots:6.1%}")
print("top lines by samples:")
for n, s, f, l, src in sorted(data, key=lambda d: -d[1])[:25]:
    print(f"{n/nrows:8.0f} smp {s/tots:6.1%} {f[:14]:14s} L{l:4d} {src}")
