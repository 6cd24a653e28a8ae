"""bench.py prints ONE JSON line with the keys the driver reads.  CPU: the reference arm (--impl reference, runs the
unmodified reference on the host cores).  GPU: the product arm on a reduced ensemble (same code path, fewer instances)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "RK4 instance-steps/s (FP64)" and d["unit"] == "instance-steps/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_product_arm_prints_the_contract_line():
    n = 1 << 16
    d = _run("--no-extras", "--steps", "2", "--warmup", "3", "--instances", str(n))
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["gpu_launches"] == 2 and d["value"] > 1e9 and d["ms_per_step"] > 0
    assert d["config"]["instances_per_gpu"] == n and "workload" in d["config"]
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.05 and e["h2d_bytes_per_step"] == 9 * 8 * n and e["d2h_bytes_per_step"] == 4 * 8 * n
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["unit"] == "TFLOP/s" and r["bound"] == "fp64"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] < 1
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_roofline_counters_are_tied_to_the_built_kernel(tmp_path, monkeypatch):
    """bench.py reports ncu counters (DRAM traffic, FP64 instructions per step) only while the SASS of the headline kernel in the
    library it loads is the one the counters were captured on; any other hash yields None plus the reason."""
    sys.path.insert(0, ROOT)
    import bench
    from sopot_b200 import build
    build.build()
    c, why = bench.load_counters(bench.N_PER_GPU)
    committed = json.load(open(bench.COUNTERS_FILE))
    if c is not None:
        assert c["sass_sha256"] == committed["sass_sha256"] and c["traffic_valid"] and c["fp64_instr_per_unit"] > 50
        assert bench.load_counters(1 << 10)[0]["traffic_valid"] is False       # another batch size: no traffic figure
    else:
        assert "stale" in why
    stale = dict(committed, sass_sha256="0" * 64)
    f = tmp_path / "c.json"
    f.write_text(json.dumps(stale))
    monkeypatch.setattr(bench, "COUNTERS_FILE", str(f))
    c, why = bench.load_counters(bench.N_PER_GPU)
    assert c is None and "stale" in why
