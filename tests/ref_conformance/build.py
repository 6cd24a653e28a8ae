"""Compile the REFERENCE'S OWN test programs, unmodified, against this repository's headers.

    python tests/ref_conformance/build.py            (run where /root/reference exists; no GPU needed)

The reference's tests include its headers by relative path ("../core/solver.hpp") or from the source root
("physics/control/lqr.hpp").  For each program an overlay source root is assembled in a temporary directory, outside the
repository:
    <tmp>/<top>/...    every header of sopot_b200/include/sopot/ (symlinks)
                       + headers that exist ONLY in the reference (units, symbolic constraints, ...: out of scope here,
                         SURVEY.md section 2) as symlinks to /root/reference, so that an in-scope test that also touches
                         an out-of-scope feature still builds; they are listed in the manifest
    <tmp>/tests/X.cpp  a copy of the reference test (never stored in the repository)
Programs whose component packs need a kernel instantiated in the test's own translation unit (N-link plants) are compiled
by nvcc -x cu, the others by g++.  Outputs: tests/ref_conformance/bin/<name> (git-ignored; they travel to the GPU box with
the snapshot like the other built artefacts) and bin/manifest.json.  tests/test_ref_conformance.py runs them on the GPU."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SOPOT_REFERENCE", "/root/reference")
INC = os.path.join(ROOT, "sopot_b200", "include", "sopot")
LIBDIR = os.path.join(ROOT, "sopot_b200", "lib")
BIN = os.path.join(HERE, "bin")

# name -> (compiler, expectation).  "runs": builds unmodified and runs on the GPU.  "rejected": the program defines its OWN
# host-only components (plain derivatives() without SOPOT_HD) or uses the reference's Cartesian / symbolic pendulums (out of
# scope); such a pack has no kernel, and the layer must refuse it at COMPILE time with the no-host-fallback message rather
# than integrate on the CPU -- the build failing with exactly that message is the expected result (their in-scope checks
# are mirrored in tests/cpp/: linearization_test.cpp, double_pendulum_test.cpp).
NO_FALLBACK = "sopot-b200 has no host fallback"
PROGRAMS = {
    "compile_time_test": ("g++", "runs"),
    "inverted_pendulum_test": ("nvcc", "runs"),
    "autodiff_test": ("g++", "rejected"),
    "double_pendulum_test": ("g++", "rejected"),
}


def overlay(tmp: str) -> list[str]:
    """Source root with this repository's headers first and reference-only files behind them; returns the latter."""
    from_ref = []
    for dirpath, _, files in os.walk(INC):
        rel = os.path.relpath(dirpath, INC)
        os.makedirs(os.path.join(tmp, rel), exist_ok=True)
        for f in files:
            os.symlink(os.path.join(dirpath, f), os.path.join(tmp, rel, f))
    for top in ("core", "physics", "rocket", "io"):
        for dirpath, _, files in os.walk(os.path.join(REF, top)):
            rel = os.path.relpath(dirpath, REF)
            os.makedirs(os.path.join(tmp, rel), exist_ok=True)
            for f in files:
                dst = os.path.join(tmp, rel, f)
                if not os.path.lexists(dst):
                    os.symlink(os.path.join(dirpath, f), dst)
                    from_ref.append(os.path.join(rel, f))
    return sorted(from_ref)


def build_one(name: str, compiler: str, expect: str, tmp: str) -> dict:
    src = os.path.join(tmp, "tests", name + ".cpp")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "tests", name + ".cpp"), src)
    out = os.path.join(BIN, name)
    common = ["-I" + tmp, "-I" + os.path.join(ROOT, "include"), "-L" + LIBDIR, "-lsopot_b200"]
    if compiler == "nvcc":
        cmd = ["nvcc", "-x", "cu", "-std=c++20", "-O2", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
               "-fmad=false", "-diag-suppress", "177,550", src, "-o", out, *common, "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../../sopot_b200/lib"]
    else:
        cmd = ["g++", "-std=c++20", "-O2", src, "-o", out, *common, "-Wl,-rpath,$ORIGIN/../../../sopot_b200/lib"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    deps = subprocess.run(([cmd[0], "-x", "cu"] if compiler == "nvcc" else ["g++"]) + ["-std=c++20", "-M", "-I" + tmp, "-I" + os.path.join(ROOT, "include"), src]
                          + (["--expt-relaxed-constexpr"] if compiler == "nvcc" else []), capture_output=True, text=True).stdout
    used = sorted({os.path.relpath(os.path.realpath(t), REF) for t in deps.replace("\\\n", " ").split()
                   if t.startswith(tmp) and os.path.realpath(t).startswith(REF + os.sep) and not t.endswith(".cpp")})
    log = r.stdout + r.stderr
    if expect == "runs":
        ok = r.returncode == 0
    else:
        ok = r.returncode != 0 and NO_FALLBACK in log
        if os.path.exists(out):
            os.remove(out)
    return {"name": name, "compiler": compiler, "expect": expect, "ok": ok, "reference_only_headers_used": used,
            "log": "" if ok else log[-3000:]}


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "tests")):
        print("reference tree not present: nothing to build")
        return 0
    os.makedirs(BIN, exist_ok=True)
    results = []
    for name, (compiler, expect) in PROGRAMS.items():
        with tempfile.TemporaryDirectory(prefix="sopot_conf_") as top:
            # the overlay mirrors the repository layout (<top>/sopot_b200/include/sopot + <top>/include), because
            # b200/engine.hpp reaches the C header by a path relative to itself
            tmp = os.path.join(top, "sopot_b200", "include", "sopot")
            os.makedirs(tmp)
            os.symlink(os.path.join(ROOT, "include"), os.path.join(top, "include"))
            overlay(tmp)
            results.append(build_one(name, compiler, expect, tmp))
            print(name, expect, "ok" if results[-1]["ok"] else "FAILED\n" + results[-1]["log"], flush=True)
    json.dump(results, open(os.path.join(BIN, "manifest.json"), "w"), indent=1)
    return 0 if all(r["ok"] for r in results) else 1


if __name__ == "__main__":
    sys.exit(main())
