"""One-page text summary of an ncu report for profiles/:  python tools/ncu_summary.py X.ncu-rep [units_per_launch] > profiles/NAME.txt
Key raw metrics (duration, DRAM bytes, pipe utilisation, occupancy, stall reasons) and the dynamic opcode mix from the
source page (per `units` work items when given)."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[-1]
m = {k: (x, uu) for k, uu, x in zip(h, u, v)}
print("report:", rep.split("/")[-1])
print("kernel:", m.get("Kernel Name", ("?",))[0][:150])
print("grid", m.get("Grid Size", ("?",))[0], "block", m.get("Block Size", ("?",))[0])
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for k in KEYS:
    if k in m:
        print(f"  {k:68s} {m[k][0]:>18s} {m[k][1]}")
print("  stall reasons (warps per issue-active cycle):")
for k in sorted(m):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            x = float(m[k][0])
        except ValueError:
            continue
        if x >= 0.1:
            print(f"    {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {x:6.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
except StopIteration:
    sys.exit(0)
hh = rows[hdr]
si, ei = hh.index("Source"), hh.index("Instructions Executed")
agg = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= ei or not r[ei]:
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    if op == "IMAD" and ".MOV" in r[si]:
        op = "IMAD.MOV"
    agg[op] += int(r[ei])
tot = sum(agg.values())
print("dynamic opcode mix: %d warp instructions" % tot + ((" = %.1f thread-instructions per unit (%g units)" % (tot * 32 / units, units)) if units else ""))
fp64 = sum(c for o, c in agg.items() if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("  FP64-pipe instructions: %.1f%% of all" % (100 * fp64 / tot) + ((", %.1f per unit" % (fp64 * 32 / units)) if units else ""))
for op, c in agg.most_common(18):
    print(f"  {c:14d} {100 * c / tot:6.2f}%  {op:10s}" + (f" {c * 32 / units:9.1f}/unit" if units else ""))
