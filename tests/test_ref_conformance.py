"""The reference's OWN test programs, unmodified, against this repository's headers (tests/ref_conformance/build.py).

CPU (where /root/reference exists): the programs whose packs have kernels build -- compile_time_test.cpp with g++,
inverted_pendulum_test.cpp (6-link plant, CartDoublePendulum cross-check, LQR<14,1>, closed loop) with nvcc -- and the
programs that define host-only components are REFUSED at compile time with the no-host-fallback message.
GPU: the built programs run and pass (the binaries travel with the snapshot; the reference tree does not)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

HERE = os.path.join(ROOT, "tests", "ref_conformance")
MANIFEST = os.path.join(HERE, "bin", "manifest.json")


def test_reference_tests_build_against_the_repo_headers():
    if not os.path.isdir("/root/reference/tests"):
        pytest.skip("reference tree not present on this machine")
    from sopot_b200 import build
    build.build()
    r = subprocess.run([sys.executable, os.path.join(HERE, "build.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    res = {x["name"]: x for x in json.load(open(MANIFEST))}
    assert res["compile_time_test"]["ok"] and res["inverted_pendulum_test"]["ok"]
    assert res["compile_time_test"]["reference_only_headers_used"] == [] and res["inverted_pendulum_test"]["reference_only_headers_used"] == []
    assert res["autodiff_test"]["expect"] == "rejected" and res["autodiff_test"]["ok"]
    assert res["double_pendulum_test"]["expect"] == "rejected" and res["double_pendulum_test"]["ok"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,needle", [("compile_time_test", "passed"), ("inverted_pendulum_test", "All cart-6-pendulum tests passed!")])
def test_reference_test_programs_pass_on_the_gpu(name, needle):
    exe = os.path.join(HERE, "bin", name)
    if not os.path.exists(exe):
        pytest.skip("tests/ref_conformance/bin/%s was not built (needs the reference tree at build time)" % name)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert needle.lower() in r.stdout.lower(), r.stdout[-2000:]
