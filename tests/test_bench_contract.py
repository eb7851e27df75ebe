"""bench.py's reference arm (the CPU path of the reference, timed on the host cores) runs without a GPU: check the
JSON contract of its line, and that under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "512", "--ny", "512",
                           "--steps", "2", "--warmup", "1", *args], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_line():
    r = run_bench()
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "MLUPS" and line["unit"] == "MLUPS"
    assert line["higher_is_better"] is True and line["value"] > 1 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"] and "sample" in base
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1 and line["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    r = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_our_arm_line_on_a_gpu():
    """The JSON contract of the product arm (small override grid so the test is quick): roofline, e2e, cpu_baseline,
    clocks and gpu_launches are all present and consistent."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--nx", "2048", "--ny", "2048", "--steps", "20", "--warmup", "3",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["metric"] == "MLUPS" and j["unit"] == "MLUPS" and j["n_gpus"] == 1 and j["steps"] == 20 and j["warmup"] == 3
    assert j["dtype"] == "f32" and j["data"] == "synthetic" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["value"] > 1000 and j["ms_per_step"] > 0 and "workload" in j["config"]
    roof = j["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["peak"] > 1000
    assert roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"], rel=1e-3)
    assert roof["achieved"] == pytest.approx(j["value"] * 1e6 * 72 / 1e9, rel=1e-3)  # 72 B per cell update
    e2e = j["e2e"]
    assert e2e["unit"] == "MLUPS" and 0 < e2e["value"] < j["value"]
    assert e2e["h2d_bytes_per_step"] == (2048 * 2048 * 37) // 20 and e2e["d2h_bytes_per_step"] > 0
    assert j["gpu_launches"] >= 20
    base = j["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] > 0
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
