"""SHA-256 of the SASS of one kernel inside a built library, and its static FP64 instruction counts.

    python tools/kernel_hash.py [lib] [mangled kernel name]      -> one JSON line

The hash ties a set of ncu counters (profiles/*_counters.json) to the code that produced them: bench.py recomputes it
from the library it loaded and refuses counters whose hash differs.  Instruction text only (addresses, encodings and
comments are stripped), so relinking the same objects keeps the hash."""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "sopot_b200", "lib", "libsopot_b200.so")
DPEND_KERNEL = "_Z19rk4_ensemble_kernelI8DpendSysLb0EEvNT_9DevParamsEmd10StepConstsmmPdS4_Pi"


def kernel_sass(lib: str, mangled: str) -> list[str]:
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        raise RuntimeError("cuobjdump not found")
    out = subprocess.run([exe, "-sass", "-fun", mangled, lib], capture_output=True, text=True, timeout=120).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            ins.append(re.sub(r"\s+", " ", m.group(1).strip()))
    if not ins:
        raise RuntimeError(f"kernel {mangled} not found in {lib}")
    return ins


def kernel_hash(lib: str = DEFAULT_LIB, mangled: str = DPEND_KERNEL) -> dict:
    ins = kernel_sass(lib, mangled)
    ops = [i.split()[1] if i.startswith("@") else i.split()[0] for i in ins]
    count = lambda p: sum(1 for o in ops if o.split(".")[0] == p)  # noqa: E731
    return {"kernel": mangled, "sass_sha256": hashlib.sha256("\n".join(ins).encode()).hexdigest(), "sass_instructions": len(ins),
            "static_dfma": count("DFMA"), "static_dmul": count("DMUL"), "static_dadd": count("DADD"), "static_mufu": count("MUFU")}


if __name__ == "__main__":
    print(json.dumps(kernel_hash(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_LIB,
                                 sys.argv[2] if len(sys.argv) > 2 else DPEND_KERNEL)))
