"""CPU: the committed bench lines (profiles/r1_bench_*.json, written by bench.py on a B200) carry every key of the
measurement contract, and bench.py's argument surface is the one the driver calls."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"}


def load(name):
    path = os.path.join(PROFILES, name)
    if not os.path.exists(path):
        pytest.skip("%s not committed" % name)
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r1_bench_ours.json", "r1_bench_c3.json", "r1_bench_c4.json", "r1_bench_c5mb.json",
                                  "r1_bench_c2_2gpu.json"])
def test_bench_line_has_contract_keys(name):
    d = load(name)
    missing = BASE_KEYS - set(d)
    if d["n_gpus"] > 1:
        missing -= {"cpu_baseline"}          # rank 0 at N=1 only
    assert not missing, missing
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] >= d["steps"]
    assert "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


def test_reference_arm_line():
    d = load("r1_bench_ref.json")
    assert d["impl"] == "reference" and d["unit"] == "lattices/s"
    ours = load("r1_bench_ours.json")
    assert d["config"]["workload"] == ours["config"]["workload"] and d["metric"] == ours["metric"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
