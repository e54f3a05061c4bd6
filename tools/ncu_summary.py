"""Summarise an .ncu-rep of one kernel: headline metrics, opcode histogram, stall reasons.
usage: python tools/ncu_summary.py report.ncu-rep rows_per_launch"""
import csv, io, subprocess, sys
from collections import Counter
rep, rows_per = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, vals = r[0], r[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {r[1][i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
s = list(csv.reader(io.StringIO(src)))
h = s[1]; body = s[2:]
ia, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
tot = sum(int(x[ia]) for x in body if x[ia].isdigit())
c = Counter(); smp = Counter()
for x in body:
    if not x[ia].isdigit():
        continue
    t = x[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(x[ia]); smp[op] += int(x[ismp]) if x[ismp].isdigit() else 0
print(f"\nwarp-instructions per row: {tot / rows_per:.0f}")
for op, nn in c.most_common(24):
    print(f"  {op:8s} {nn / tot * 100:5.1f}%  per-row {nn / rows_per:8.1f}  samples {smp[op]}")
st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
agg = Counter()
for x in body:
    for i in st:
        try:
            agg[h[i]] += int(x[i])
        except ValueError:
            pass
ss = sum(agg.values())
print("\nstall samples:")
for k, v in agg.most_common(10):
    print(f"  {k:26s} {v / ss * 100:5.1f}%")
