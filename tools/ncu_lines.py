"""Stall samples and executed instructions of one kernel aggregated by CUDA source line:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv ; python tools/ncu_lines.py cs.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
samp, inst, txt = collections.Counter(), collections.Counter(), {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] == "Line No" or r[2] != "-":
        continue
    try:
        ln, s, n = int(r[0]), int(r[4] or 0), int(r[7] or 0)
    except ValueError:
        continue
    samp[(cur, ln)] += s
    inst[(cur, ln)] += n
    txt[(cur, ln)] = r[1].strip()[:100]
ts, ti = sum(samp.values()), sum(inst.values())
print(f"{ts} stall samples, {ti} warp instructions")
print("samples%  instr%  line")
for k, v in samp.most_common(top):
    print(f"{100 * v / ts:6.1f}  {100 * inst[k] / ti:6.1f}   {k[0]}:{k[1]}  {txt[k]}")
