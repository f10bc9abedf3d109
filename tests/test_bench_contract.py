"""CPU: the committed bench lines (profiles/r1_bench_*.json and r2_bench_*.json, written by bench.py on a B200) carry
every key of the measurement contract, and bench.py's argument surface is the one the driver calls."""
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


R2 = ["r2_bench_ours.json", "r2_bench_c2g.json", "r2_bench_c3.json", "r2_bench_c3d.json", "r2_bench_c4.json",
      "r2_bench_c4d.json", "r2_bench_c5mb.json", "r2_bench_c2b.json", "r2_bench_c5mbb.json", "r2_bench_c2l.json",
      "r2_bench_c5mbl.json", "r2_bench_c2_2gpu_builder.json", "r2_bench_c2_4gpu_builder.json",
      "r2_bench_c2_8gpu_builder.json"]


@pytest.mark.parametrize("name", R2)
def test_round2_bench_line_has_contract_keys(name):
    d = load(name)
    missing = BASE_KEYS - set(d)
    if d["n_gpus"] > 1 or name != "r2_bench_ours.json":
        missing -= {"cpu_baseline"}          # rank 0 at N=1 only; the extra workloads were run with --no-cpu-baseline or with it
    assert not missing, missing
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] in ("f32", "bf16 i/o, f32 accumulate")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] >= d["steps"]
    assert "workload" in d["config"] and "timed_call" in d
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and r["kernel_ms"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] > 1:
        c5 = d["c5"]
        assert c5["value"] > 0 and 0 < c5["roofline_frac_per_gpu"] < 1 and "N=2048" in c5["workload"]


def test_round2_both_arms_share_the_config_dict():
    ours, ref = load("r2_bench_ours.json"), load("r2_bench_ref.json")
    assert ref["impl"] == "reference" and ours["config"] == ref["config"] and ours["metric"] == ref["metric"]
    assert set(ref["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}


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
