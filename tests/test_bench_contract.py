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
