"""Static instruction count of the hot path through the time loop of a kernel, from its SASS alone (no GPU).

    python tools/sass_hot_path.py [lib] [mangled kernel name]      -> one JSON line

The time loop of the ensemble kernels is a DAG of basic blocks between the loop head and the backward branch; every
shortcut of a step (small / tiny / micro rotations, no refresh) is the cheaper side of its branch, so the path with the
fewest FP64 instructions is the one an instance takes on an ordinary step.  Printed per iteration of that path: FP64-pipe
instructions by opcode, DFMAs reading three distinct vector-register pairs (3 issue cycles on sm_100a instead of 2,
profiles/r1_fp64_operand_rule.txt; how many of them carry a .reuse flag is listed beside it), all instructions, and
the FP64 issue cycles that rule predicts.  Checked against ncu on the round-2 kernel: 174.0 executed FP64 instructions and 45.5
three-register DFMAs per instance-step measured, 172 + 45 on the static hot path (the rest is the 1-in-16 refresh)."""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_hash import DEFAULT_LIB, DPEND_KERNEL  # noqa: E402
import shutil  # noqa: E402
import subprocess  # noqa: E402

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def disasm(lib, mangled):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", "-fun", mangled, lib], capture_output=True, text=True, timeout=120).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r"\s+", " ", m.group(2).strip())))
    if not ins:
        raise RuntimeError(f"kernel {mangled} not found in {lib}")
    return ins


def opcode(text):
    t = text.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


def vec_sources(text):
    """distinct vector-register pairs among the sources of an FP64 instruction, and whether a source carries .reuse"""
    body = text.split(None, 2 if text.startswith("@") else 1)[-1]
    ops = [o.strip() for o in body.split(",")]
    srcs = ops[1:]
    regs = set()
    reuse = False
    for s in srcs:
        m = re.match(r"[-|~]*\|?(R\d+)", s)
        if m and m.group(1) != "RZ":
            regs.add(m.group(1))
            reuse |= ".reuse" in s
    return len(regs), reuse


def hot_path(lib=DEFAULT_LIB, mangled=DPEND_KERNEL):
    ins = disasm(lib, mangled)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    # the time loop: the LAST backward branch whose target is the lowest (outermost loop with the largest body)
    back = []
    for i, (a, t) in enumerate(ins):
        if opcode(t) == "BRA":
            m = re.search(r"(0x[0-9a-f]+)\s*$", t)
            if m and int(m.group(1), 16) < a:
                back.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
    if not back:
        raise RuntimeError("no loop found")
    _, head, tail = max(back)
    lo, hi = addr_index[head], addr_index[tail]
    # successor lists inside [lo, hi]
    succ = {}
    for i in range(lo, hi + 1):
        a, t = ins[i]
        op = opcode(t)
        nxt = []
        if op == "BRA":
            m = re.search(r"(0x[0-9a-f]+)\s*$", t)
            tgt = int(m.group(1), 16)
            cond = t.startswith("@") or ".U" in t.split()[0] or re.search(r"BRA(\.\w+)*\s+!?U?P\d", t) is not None
            if i == hi:
                nxt = []                      # the back edge ends the iteration
            else:
                if tgt in addr_index and lo <= addr_index[tgt] <= hi:
                    nxt.append(addr_index[tgt])
                if cond:
                    nxt.append(i + 1)
        elif op in ("EXIT", "RET"):
            nxt = [] if not t.startswith("@") else [i + 1]
        else:
            nxt = [i + 1] if i < hi else []
        succ[i] = nxt
    # dynamic programming from the tail backwards over the DAG (forward edges only)
    INF = (1 << 60,)
    best = {}

    def cost(i):
        _, t = ins[i]
        op = opcode(t)
        if op == "CALL":
            return 10000                      # an out-of-line routine is never on the hot path (its instructions are not in this listing)
        return 1 if op in FP64 else 0

    order = range(hi, lo - 1, -1)
    for i in order:
        c = cost(i)
        nx = [j for j in succ[i] if j > i]
        if i == hi:
            best[i] = (c, [i])
            continue
        if not nx:
            best[i] = (INF[0], [i])           # leaves the loop (exit path): never the hot path
            continue
        j = min(nx, key=lambda j: best[j][0])
        best[i] = (c + best[j][0], None if best[j][0] >= INF[0] else j)
    path = []
    i = lo
    while True:
        path.append(i)
        if i == hi:
            break
        nxt = best[i][1]
        if nxt is None:
            raise RuntimeError("hot path leaves the loop")
        i = nxt
    counts = {}
    three = three_reuse = 0
    for i in path:
        _, t = ins[i]
        op = opcode(t)
        counts[op] = counts.get(op, 0) + 1
        if op == "DFMA":
            n, reuse = vec_sources(t)
            if n >= 3:
                three += 1
                three_reuse += reuse
    fp64 = sum(counts.get(o, 0) for o in FP64)
    return {"kernel": mangled, "loop": [hex(head), hex(tail)], "path_instructions": len(path), "fp64": fp64,
            "dfma": counts.get("DFMA", 0), "dmul": counts.get("DMUL", 0), "dadd": counts.get("DADD", 0),
            "dsetp": counts.get("DSETP", 0), "mufu": counts.get("MUFU", 0), "dfma_three_vector_sources": three,
            "of_which_with_reuse_flag": three_reuse, "fp64_issue_cycles_by_rule": 2 * fp64 + three,
            "other": {k: v for k, v in sorted(counts.items()) if k not in FP64 and k != "MUFU"}}


if __name__ == "__main__":
    print(json.dumps(hot_path(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_LIB,
                              sys.argv[2] if len(sys.argv) > 2 else DPEND_KERNEL)))
