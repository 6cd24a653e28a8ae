"""CPU tests of the drop-in boundary: libsopot_b200.so loads, exports exactly what include/sopot_b200.h
declares, and refuses to run (loudly) without a GPU -- there is no CPU fallback to fall into."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "sopot_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sopot_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from sopot_b200 import build, capi
    lib_path = build.build()
    assert os.path.exists(lib_path)
    declared = _declared()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (sopot_b200_[a-z0-9_]+)", out)))
    assert exported == declared, (set(declared) ^ set(exported))
    assert sorted(capi.SYMBOLS) == declared            # the ctypes table binds all of them
    capi.load_library()


def test_library_carries_sm100a_code_only():
    from sopot_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    from sopot_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SopotB200Error, match="no CPU fallback"):
        capi.Engine(0)


def test_product_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sopot_b200")):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(import|from)\s+oracle|oracle/|libsopot_oracle|libsopot_ref|orc_[a-z]", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_num_steps_rule_matches_reference_loop():
    from sopot_b200 import capi
    lib = capi.load_library()
    assert lib.sopot_b200_rk4_num_steps(0.0, 10.0, 0.01) == 1000
    assert lib.sopot_b200_rk4_num_steps(0.0, 10.0, 1e-3) == 10000
    assert lib.sopot_b200_rk4_num_steps(0.0, 30.0, 0.01) == 3000
    assert lib.sopot_b200_rk4_num_steps(0.0, 0.016, 0.01) == 2
    assert lib.sopot_b200_rk4_num_steps(0.0, 1.0, 0.0) == 0


def test_headline_kernel_counters_belong_to_the_built_code():
    """profiles/dpend_bench_counters.json (the ncu counters bench.py reports as roofline.traffic / fp64_instr_per_unit) was
    captured on exactly the kernel this tree builds -- same SHA-256 of its SASS -- and agrees with the static hot path of
    that kernel's time loop (tools/sass_hot_path.py: the executed count exceeds the static one only by the 1-in-16 refresh
    and the rare fall-back tiers; the three-register DFMA count likewise).  No GPU needed: cuobjdump on the built library."""
    import json
    import sys
    from sopot_b200 import build
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from kernel_hash import kernel_hash
    from sass_hot_path import hot_path
    lib = build.build()
    c = json.load(open(os.path.join(ROOT, "profiles", "dpend_bench_counters.json")))
    assert kernel_hash(lib, c["kernel_mangled"])["sass_sha256"] == c["sass_sha256"], \
        "the double-pendulum kernel changed: re-capture profiles/dpend_bench_counters.json (tools/ncu_counters.py)"
    h = hot_path(lib, c["kernel_mangled"])
    assert h["fp64"] <= c["fp64_instr_per_unit"] <= h["fp64"] + 4, (h["fp64"], c["fp64_instr_per_unit"])
    assert abs(h["dfma_three_vector_sources"] - c["dfma3_per_unit"]) <= 2, (h["dfma_three_vector_sources"], c["dfma3_per_unit"])
    assert h["mufu"] == 4 and h["dsetp"] == 0           # four reciprocal seeds, every range test on the integer pipe
    assert h["fp64"] <= 156                             # the third form of the step (csrc/dpend_step.cuh); the second had 172
