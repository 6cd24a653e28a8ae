"""SASS listings of the hot kernels for profiles/sass/ (instruction text only, one file per kernel):
    python tools/dump_sass.py
Kernels are found by a substring of their mangled name in the built library."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sopot_b200", "lib", "libsopot_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
WANT = {
    "dpend_contract": "rk4_ensemble_kernelI8DpendSysLb0",
    "grid_march_contract": "grid_rk4_marchILb0ELb0",
    "grid_march_contract_p2p": "grid_rk4_marchILb0ELb1",        # several ranks: exchange over NVLink peer memory inside the kernel
    "grid_march_strict": "grid_rk4_marchILb1ELb0",
    "ugrid_march_strict_stage2": "ugrid_stage_marchILb1ELi2",
    "rocket_contract": "rk4_ensemble_kernelI10RocketSysTILb0EELb0",
    "ho_contract": "rk4_ensemble_kernelI5HoSysLb0",
    "cartpole_linearize_contract": "cartpole_linearize_kernelILb0",
}
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
blocks, name = {}, None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        blocks[name] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and name:
        blocks[name].append("/*%s*/  %s" % (m.group(1), re.sub(r"\s+", " ", m.group(2).strip())))
os.makedirs(OUT, exist_ok=True)
for tag, key in WANT.items():
    hits = [n for n in blocks if key in n]
    if not hits:
        print("missing", tag, file=sys.stderr)
        continue
    n = hits[0]
    ops = {}
    for l in blocks[n]:
        t = l.split()[1:]
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    head = ["// %s" % n, "// %d instructions; static mix: %s" % (len(blocks[n]), ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))]
    open(os.path.join(OUT, tag + ".sass"), "w").write("\n".join(head + blocks[n]) + "\n")
    print(tag, len(blocks[n]))
