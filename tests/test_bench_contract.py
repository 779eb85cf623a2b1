"""The bench.py JSON-line contract, checked on the lines recorded from real B200 runs
(profiles/raw) and on the argument parsing / reference arm plumbing that runs without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = os.path.join(ROOT, "profiles", "raw")


def _line(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("fname,n_gpus", [("r01_bench_1gpu.json", 1), ("r01_bench_2gpu.json", 2),
                                          ("r01_bench_4gpu.json", 4), ("r01_bench_8gpu.json", 8),
                                          ("r01b_bench_1gpu.json", 1), ("r01b_bench_2gpu.json", 2)])
def test_recorded_bench_lines_follow_the_contract(fname, n_gpus):
    d = _line(os.path.join(RAW, fname))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == n_gpus and d["metric"] == "effective_hbm_gbps_per_gate" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["d2h_bytes_per_step"] > 0 and e["h2d_bytes_per_step"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n_gpus == 1:
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] == 1 and c["value"] > 0 and "sample" in c
        assert e["d2h_bytes_per_step"] >= d["config"]["state_bytes"]
    else:
        assert d["nvlink"]["nvlink_bytes_sent_per_gpu_per_step"] > 0


def test_reference_arm_line():
    d = _line(os.path.join(RAW, "r01_bench_reference_arm.json"))
    assert d["impl"] == "reference" and d["metric"] == "effective_hbm_gbps_per_gate"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
