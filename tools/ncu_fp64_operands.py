"""FP64-pipe instructions of one kernel by number of DISTINCT VECTOR source registers, from an ncu source page:
   ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_fp64_operands.py src.csv <work items>
A DFMA reading three distinct vector registers issues every 3 cycles on sm_100a, anything else every 2
(profiles/r1_fp64_operand_rule.txt); the last line is the resulting FP64 issue cycles per warp and work item."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) / 32.0
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
si, ei = h.index("Source"), h.index("Instructions Executed")
cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= ei or not r[ei]:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?(D(?:FMA|MUL|ADD|SETP))\S*\s+(.*)", r[si])
    if not m:
        continue
    regs = set()
    for o in m.group(3).split(",")[1:]:
        mm = re.match(r"R(\d+)", o.strip().lstrip("-|"))
        if mm:
            regs.add(mm.group(1))
    cnt[(m.group(2), len(regs))] += int(r[ei])
tot = cyc = 0.0
for k, v in sorted(cnt.items()):
    print(f"{k[0]:6s} {k[1]} vector sources: {v / units:8.2f} per item")
    tot += v / units
    cyc += v / units * (3 if k == ("DFMA", 3) else 2)
print(f"FP64 instructions per item {tot:.1f}; issue cycles per warp-item {cyc:.1f}")
