"""Static SASS opcode histogram of libsakura_b200.so: cuobjdump -sass sakura_b200/libsakura_b200.so | python tools/sass_histogram.py"""
import re
import subprocess
import sys
from collections import Counter, defaultdict

fn = None
per = defaultdict(Counter)
tot = Counter()
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and fn:
        op = m.group(2).split('.')[0]
        per[fn][op] += 1
        tot[op] += 1
print("SASS opcode histogram of sakura_b200/libsakura_b200.so (cuobjdump -sass, sm_100a), static instruction counts")
print("\nwhole library, opcodes that show which hardware paths the code uses:")
for op in ("DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMNMX3", "REDUX",
           "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "HMMA", "IMMA"):
    print(f"  {op:8s} {tot.get(op, 0)}")
print("\n(no UTC*MMA / LDTM / STTM / UTMALDG: tcgen05 has no FP64 kind and the tensor-core work of this library is FP64 (DMMA);\n"
      " bulk asynchronous copies (UBLKCP) with mbarriers (SYNCS) stage rows and table tiles, LDGSTS = cp.async; DESIGN.md 3.1c, 3.2)")
print("\nper kernel (largest first), top opcodes:")
for f, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:48]:
    try:
        name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    except Exception:
        name = f
    name = re.sub(r'sakura_b200::\(anonymous namespace\)::', '', name)
    print(f"\n{name[:140]}\n  {sum(c.values())} instructions: " + ", ".join(f"{o} {k}" for o, k in c.most_common(12)))
