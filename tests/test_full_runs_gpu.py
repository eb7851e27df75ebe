"""Whole jobs on the GPU: the reference's four shipped configurations for their full maxIters through the
lbm_runner driver and through the Python binding, validated like the reference validates itself (check.py
against check/*.dat) and, more strictly, against the oracle fixtures: the final lattice must hash to the oracle's
SHA-256 after 20 000 - 80 000 timesteps.  Then BASELINE.json's full size (16384 x 16384) through size-independent
properties and band-exact comparison with the oracle."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import CASES, golden_av_vels, golden_final_state, obstacles_path, oracle_fixture, params_path, ulp_distance
from tools import check as check_tool

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES)
def test_lbm_runner_full_run_passes_check(lbm, case, tmp_path):
    r = subprocess.run([lbm.RUNNER_PATH, "--params", params_path(case), "--obstacles", obstacles_path(case), "--outdir",
                        str(tmp_path), "--reruns", "0", "-d"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "==done==" in r.stdout and "Total compute time was" in r.stdout and "Running LBM" in r.stderr
    fx = oracle_fixture(case)
    # the driver prints reynoldsNumber(av_vels[maxIters-1]) like main/LbmRunner.cpp:129-130 (main/LastChance.cpp
    # prints the Reynolds number of the FINAL STATE instead -- that one is fx["reynolds"], 1e-3 away)
    p = lbm.Params.from_file(params_path(case))
    re_num = float(re.search(r"Reynolds number:\s+(\S+)", r.stdout).group(1))
    assert re_num == pytest.approx(p.reynolds_number(float(fx["av_f64sum"][-1])), rel=2e-6)
    assert re_num == pytest.approx(float(fx["reynolds"]), rel=5e-3)
    av = np.loadtxt(tmp_path / "av_vels.dat", usecols=[1])
    fs = np.loadtxt(tmp_path / "final_state.dat", usecols=[0, 1, 5])
    # the reference's own acceptance rule (1 %) against its shipped double-precision goldens
    d = check_tool.get_diff_values(golden_av_vels(case), av)
    assert np.isfinite(d["max_diff_pcnt"]) and abs(d["max_diff_pcnt"]) < 1.0, d
    if case in ("128x128", "128x256"):
        code, _, _ = check_tool.compare(golden_av_vels(case), golden_final_state(case), av, fs, 1.0, out=open(os.devnull, "w"))
        assert code == 0
    # and the oracle: av_vels to 1 fp32 ulp, pressure field exactly
    assert ulp_distance(av.astype(np.float32), fx["av_f64sum"]).max() <= 1
    s = int(fx["stride"])
    pressure = fs[:, 2].astype(np.float32).reshape(p.ny, p.nx)[::s, ::s]
    assert np.array_equal(pressure, fx["pressure"])
    if case == "128x128":
        # the reference's way (whole lattice back, fields on the host) writes the same bytes
        other = tmp_path / "cells_way"
        other.mkdir()
        r2 = subprocess.run([lbm.RUNNER_PATH, "--params", params_path(case), "--obstacles", obstacles_path(case), "--outdir",
                             str(other), "--reruns", "0", "--download-cells"], capture_output=True, text=True, timeout=900)
        assert r2.returncode == 0
        assert (other / "final_state.dat").read_bytes() == (tmp_path / "final_state.dat").read_bytes()
        assert (other / "av_vels.dat").read_bytes() == (tmp_path / "av_vels.dat").read_bytes()


@pytest.mark.parametrize("case", CASES)
def test_full_run_lattice_hash_equals_oracle(lbm, case):
    p = lbm.Params.from_file(params_path(case))
    obst = lbm.load_obstacles(p.nx, p.ny, obstacles_path(case))
    fx = oracle_fixture(case)
    with lbm.Solver(p) as s:
        s.upload(lbm.initial_cells(p), obst)
        s.run(p.max_iters)
        cells, av = s.download()
    assert hashlib.sha256(cells.tobytes()).hexdigest() == str(fx["cells_sha256"])
    assert np.array_equal(cells[p.ny - 2].view(np.uint32), fx["cells_row"].view(np.uint32))  # carries the extra accelerate (F8)
    assert ulp_distance(av, fx["av_f64sum"]).max() <= 1


def test_baseline_size_properties_and_band_parity(oracle, lbm):
    """16384 x 16384 synthetic channel (BASELINE.json configs[3]).  (1) mass is conserved; (2) two runs give the
    same tot_u bits; (3) information travels one row per step, so K more steps of a band of rows depend only on
    the band plus K rows either side: the oracle run on that band must reproduce the GPU's full-size result BIT
    FOR BIT on the band's interior -- checked on the band under the driven lid (where the flow is) and on a band
    in the middle of the channel."""
    nx = ny = 16384
    k0, k1, seed, prob = 150, 6, 1234, 0.01
    p = lbm.Params(nx, ny, k0 + k1, 10, 0.1, 0.005, 1.85)
    with lbm.Solver(p) as s:
        s.init_synthetic(seed, prob)
        free = s.free_cells
        s.run(k0)
        before, _ = s.download(av_vels=False)
        s.run(k1)
        after, _ = s.download(av_vels=False)
        tot = s.download_tot_u()
    with lbm.Solver(p) as s:
        s.init_synthetic(seed, prob)
        s.run(k0 + k1)
        tot2 = s.download_tot_u()
    assert np.array_equal(tot.view(np.uint64), tot2.view(np.uint64))  # deterministic reduction
    assert np.all(np.isfinite(tot)) and np.all(np.diff(tot[:k0]) > 0)  # the lid keeps accelerating the flow
    mass0 = float(nx) * ny * float(np.float32(0.1))
    for lattice in (before, after):
        m = 0.0
        for r in range(0, ny, 1024):
            m += float(lattice[r:r + 1024].sum(dtype=np.float64))
        assert m == pytest.approx(mass0, rel=2e-6)
    margin = k1 + 3
    for lo, hi in ((ny - 96, ny), (ny // 2 - 40, ny // 2 + 56)):
        rows = hi - lo
        band_obst = lbm.synthetic_obstacles(nx, ny, seed, prob, lo, rows)
        assert int((band_obst == 0).sum()) > 0 and free > 0
        op = oracle.make_params(nx, rows, k1, 10, 0.1, 0.005, 1.85)
        src = np.ascontiguousarray(before[lo:hi])
        dst = np.empty_like(src)
        threads = os.cpu_count() or 1
        for _ in range(k1):
            oracle.timestep_rows(op, src, dst, band_obst, 0, rows, oracle.ACCUM_DOUBLE, threads)
            src, dst = dst, src
        # the band's own periodic wrap and its local "row ny-2" only contaminate `margin` rows at an open edge; the
        # top band ends at the solid wall row ny-1 (whose own upward-reflected populations see the wrong row above
        # it, but never send anything back down), and its local driven row IS the global one
        a = margin
        b = rows - 1 if hi == ny else rows - margin
        got = after[lo + a:lo + b]
        assert np.array_equal(got.view(np.uint32), src[a:b].view(np.uint32)), (lo, hi)
        assert hi != ny or np.abs(got[..., 1] - got[..., 3]).max() > 1e-4  # there is flow under the lid


def test_lbm_runner_synthetic_channel(lbm, tmp_path):
    """lbm_runner --synthetic NXxNY generates the channel on the device: same av_vels as the Python binding."""
    r = subprocess.run([lbm.RUNNER_PATH, "--synthetic", "512x384", "--iters", "60", "--outdir", str(tmp_path), "--reruns", "1", "-d"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MLUPS" in r.stdout and not (tmp_path / "final_state.dat").exists()
    av = np.loadtxt(tmp_path / "av_vels.dat", usecols=[1]).astype(np.float32)
    p = lbm.Params(512, 384, 60, 10, 0.1, 0.005, 1.85)
    with lbm.Solver(p) as s:
        s.init_synthetic(1234, 0.01)
        s.run(60)
        _, want = s.download(cells=False)
    assert np.array_equal(av.view(np.uint32), want.view(np.uint32))
