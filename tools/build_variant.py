"""Build an experimental variant of libsopot_b200.so: python tools/build_variant.py <tag> <file.cu> [-D...]
Recompiles ONE translation unit with extra defines and links it with the stock objects into
sopot_b200/lib/variants/libsopot_b200_<tag>.so (selected at run time with SOPOT_B200_LIB=...)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sopot_b200 import build as B  # noqa: E402

tag, unit, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
vdir = os.path.join(B.LIBDIR, "variants")
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(B.OBJ, f"{unit[:-3]}_{tag}.o")
r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + defs + ["-c", os.path.join(B.CSRC, unit), "-o", obj],
                   capture_output=True, text=True)
if r.returncode:
    raise SystemExit(r.stdout + r.stderr)
for line in (r.stdout + r.stderr).splitlines():
    if "rk4_ensemble" in line and "Lb0" in line and "Compiling" in line:
        print(line.strip()[:110])
    if "registers" in line or "spill" in line:
        last = line.strip()
open(obj + ".ptxas.log", "w").write(r.stdout + r.stderr)
# link the variant object with the stock objects of every OTHER translation unit (objects of earlier variants are skipped)
stock = {os.path.basename(src)[:-3] + ".o" for src in os.listdir(B.CSRC) if src.endswith(".cu")} - {unit[:-3] + ".o"}
objs = [os.path.join(B.OBJ, f) for f in sorted(stock)]
out = os.path.join(vdir, f"libsopot_b200_{tag}.so")
subprocess.run([B._nvcc(), "-shared", "-o", out, obj] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                  "-Xcompiler", "-fPIC", "-ldl", "-lpthread"], check=True)
print(out)
