"""Dynamic opcode mix of one kernel from an ncu report's source page:
   ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_opmix.py src.csv [units]
Prints warp-level instructions executed per opcode (and per `units` work items when given)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
si, ei = h.index("Source"), h.index("Instructions Executed")
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
print("total warp instructions", tot, ("= %.1f thread-instr per unit" % (tot * 32 / units)) if units else "")
for op, c in agg.most_common(30):
    print(f"{c:14d} {100 * c / tot:6.2f}%  {op}" + (f"   {c * 32 / units:8.1f}/unit" if units else ""))
