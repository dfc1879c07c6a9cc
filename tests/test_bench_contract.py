"""bench.py's reference arm runs on the host cores alone and prints one JSON line with the contract's keys; the
committed bench lines of our arm (profiles/) carry roofline, cpu_baseline, e2e, clocks and the launch count."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_prints_the_contract_line(oracle_ref):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-n", "12"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and BASE_KEYS <= set(line)
    assert line["metric"] == "element-steps/s" and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["value"] > 0 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("name", ["r01d_bench.json", "r01d_bench_tet.json", "r01d_bench_quad.json"])
def test_committed_bench_lines_follow_the_contract(name):
    path = os.path.join(ROOT, "profiles", name)
    line = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line)
    assert line["unit"] == "element-steps/s" and line["n_gpus"] == 1 and line["warmup"] >= 3
    assert line["gpu_launches"] == 4 * line["steps"]
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert set(roof["passes"]) == {"E1 element volume", "N1 nodal sums", "E2 main element pass", "N2 assembly+integration"}
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] < line["value"]
    assert not line["config"]["nonfinite"] and "workload" in line["config"]
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    assert not bad
