"""bench.py's reference arm (the CPU path of the reference, timed on the host cores) runs without a GPU: check the
JSON contract of its line, and that under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

import pytest

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
    # the arm runs what it prints: W + K real timesteps of the WHOLE grid named in config, ms_per_step measured
    assert "1 + 2 timesteps of the whole 512x512 grid" in base["sample"]
    assert line["detail"]["rows_timed"] == 512
    assert line["ms_per_step"] == pytest.approx(line["detail"]["timed_s"] * 1e3 / 2, rel=1e-2)
    assert line["value"] == pytest.approx(512 * 512 / (line["ms_per_step"] * 1e-3) / 1e6, rel=2e-2)
    # `config` describes the workload only (our arm prints the same object; run-specific facts live in `detail`)
    assert set(line["config"]) == {"workload", "nx", "ny", "density", "accel", "omega", "seed", "solid_prob", "l2"}


def test_both_arms_build_the_same_config_object():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = bench.workload_config("synthetic channel 16384x16384", 16384, 16384)
    assert a["l2"].startswith("working set") and "no flush" in a["l2"]
    assert bench.workload_config("x", 512, 512)["l2"].startswith("development override")


def test_reference_arm_other_ranks_stay_silent():
    r = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


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
    assert set(j["config"]) == {"workload", "nx", "ny", "density", "accel", "omega", "seed", "solid_prob", "l2"}
    assert set(j["detail"]) >= {"partition", "kernel_variant", "av_vels_last", "wall_ms_per_step"}
    par = j["parity"]  # one GPU: the multi-slab kernel on two slabs sharing the GPU, and one slab, both against the oracle
    assert par["ok"] and par["halo_lattice_sha_equal"] and par["one_slab_lattice_sha_equal"] and par["av_vels_max_ulp"] <= 1
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


@pytest.mark.gpu
def test_partition_comparison_mode_on_one_gpu():
    """bench.py --in-process N [--grid PXxPY] (the row-slab vs block comparison runs of DESIGN 5.2), here with one GPU: a 1 x 1 block
    "grid" whose ghost columns wrap onto the block itself, against the plain single slab."""
    out = {}
    for extra in ([], ["--grid", "1x1"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--in-process", "1", "--nx", "2048", "--ny", "1024", "--steps", "20",
                            "--warmup", "3", *extra], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert j["n_gpus"] == 1 and j["e2e"] is None and "parity" not in j and "shipped" not in j
        out[bool(extra)] = j
    assert "blocks with ghost columns" in out[True]["detail"]["partition"] and "row slabs" in out[False]["detail"]["partition"]
    assert out[True]["detail"]["av_vels_last"] == pytest.approx(out[False]["detail"]["av_vels_last"], rel=1e-6)
    assert out[True]["roofline"]["kernel"].endswith(",true>") and out[False]["roofline"]["kernel"].endswith(",false>")
