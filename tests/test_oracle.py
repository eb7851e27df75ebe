"""CPU tests of the oracle (oracle/): known-answer vectors of the reference's test/codelets/main.cpp, the
reference's shipped goldens (check/*.dat) under check.py's rule, and -- where oracle/_ref exists -- the
unmodified reference program itself.  Nothing here touches the GPU or the product library."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import CASES, golden_av_vels, golden_final_state, obstacles_path, oracle_fixture, params_path, ulp_distance
from tools import check as check_tool


# ---- known-answer tests ---------------------------------------------------------
def test_kat_accelerate(oracle):
    """test/codelets/main.cpp:413-480: ny=2, nx=3, density=9, accel=1 => w1=1, w2=0.25."""
    p = oracle.make_params(3, 2, 1, 1, 9.0, 1.0, 1.0)
    cells = np.array([[[1, 0.5, 1, 1, 1, 1, 1, 1, 1], [0, 1, 2, 3, 4, 5, 6, 7, 8], [0, 1, 2, 3, 4, 5, 6, 7, 8]],
                      [[2, 3, 4, 5, 6, 7, 8, 9, 10]] * 3], dtype=np.float32)
    obst = np.array([[0, 1, 0], [0, 1, 1]], dtype=np.uint8)
    want = cells.copy()
    want[0, 2] = [0, 2, 2, 2, 4, 5.25, 5.75, 6.75, 8.25]  # only the free cell that stays positive moves
    oracle.accelerate_flow(p, cells, obst)
    assert np.array_equal(cells, want)


def test_kat_rebound(oracle):
    """test/codelets/main.cpp:988-991."""
    s = np.array([2.30, 2.31, 2.32, 2.33, 2.34, 2.35, 2.36, 2.37, 2.38], dtype=np.float32)
    want = np.array([2.30, 2.33, 2.34, 2.31, 2.32, 2.37, 2.38, 2.35, 2.36], dtype=np.float32)
    assert np.array_equal(oracle.rebound_cell(s), want)


def test_kat_collision_matches_textbook_bgk(oracle):
    """test/codelets/main.cpp:879-930: omega=1, f + omega (f_eq - f) with the textbook equilibrium,
    ASSERT_FLOAT_EQ (4 ulp)."""
    f = np.float32
    s = np.array([2.30, 2.31, 2.32, 2.33, 2.34, 2.35, 2.36, 2.37, 2.38], dtype=np.float32)
    omega, c_sq = f(1.0), f(1.0) / f(3.0)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4, dtype=np.float32)
    rho = f(0)
    for v in s:
        rho = f(rho + v)
    ux = f(f(s[1] + s[5] + s[8] - s[3] - s[6] - s[7]) / rho)
    uy = f(f(s[2] + s[5] + s[6] - s[4] - s[7] - s[8]) / rho)
    u = np.array([0, ux, uy, -ux, -uy, ux + uy, -ux + uy, -ux - uy, ux - uy], dtype=np.float32)
    usq = f(ux * ux + uy * uy)
    feq = (w * rho * (f(1.0) + u / c_sq + u * u / (f(2.0) * c_sq * c_sq) - usq / (f(2.0) * c_sq))).astype(np.float32)
    want = (s + omega * (feq - s)).astype(np.float32)
    got, u_sq = oracle.collide_cell(s, 1.0)
    assert ulp_distance(got, want).max() <= 4
    assert abs(u_sq - float(usq)) <= 1e-9
    # collision conserves mass and momentum
    assert abs(float(got.astype(np.float64).sum() - s.astype(np.float64).sum())) < 1e-5


def test_kat_normed_velocity_and_masked_mean(oracle):
    """test/codelets/main.cpp:57-67, :386-403: sqrt(ux^2+uy^2) per cell, mean over non-obstacle cells."""
    cells = np.array([[[1, 2, 3, 4, 5, 6, 7, 8, 9], [0, 1, 2, 3, 4, 5, 6, 7, 8], [2, 3, 4, 5, 6, 7, 8, 9, 10]]], np.float32)

    def vel(c):
        c = c.astype(np.float32)
        rho = np.float32(0)
        for v in c:
            rho = np.float32(rho + v)
        ux = np.float32((c[1] + c[5] + c[8] - (c[3] + c[6] + c[7])) / rho)
        uy = np.float32((c[2] + c[5] + c[6] - (c[4] + c[7] + c[8])) / rho)
        return np.float32(np.sqrt(np.float32(ux * ux + uy * uy)))

    p = oracle.make_params(3, 1, 1, 1, 0.1, 0.005, 1.0)
    obst = np.array([[0, 1, 0]], dtype=np.uint8)
    want = np.float32(np.float32(vel(cells[0, 0]) + vel(cells[0, 2])) / np.float32(2))
    assert oracle.av_velocity(p, cells, obst) == pytest.approx(float(want), rel=1e-6)
    ux, uy, u, pr = oracle.final_state_fields(p, cells, obst)
    assert u[0, 0] == vel(cells[0, 0]) and u[0, 2] == vel(cells[0, 2]) and u[0, 1] == 0


def test_obstacle_cell_is_untouched_by_collision_but_rebounds(oracle):
    """test/codelets/main.cpp:872-876 + the fused step: a solid cell only swaps opposite populations."""
    p = oracle.make_params(4, 4, 1, 1, 0.1, 0.005, 1.7)
    rng = np.random.default_rng(1)
    src = (0.1 * (1 + rng.random((4, 4, 9)))).astype(np.float32)
    dst = np.zeros_like(src)
    obst = np.zeros((4, 4), np.uint8)
    obst[1, 2] = 1
    oracle.timestep_rows(p, src, dst, obst, 0, 4, oracle.ACCUM_FLOAT, 1)
    pulled = np.array([src[1, 2, 0], src[1, 1, 1], src[0, 2, 2], src[1, 3, 3], src[2, 2, 4], src[0, 1, 5], src[0, 3, 6],
                       src[2, 3, 7], src[2, 1, 8]], np.float32)
    assert np.array_equal(dst[1, 2], oracle.rebound_cell(pulled))


def test_periodic_wrap_in_both_directions(oracle):
    """main/LastChance.cpp:196-200: the grid is a torus; walls exist only as obstacle cells."""
    p = oracle.make_params(3, 2, 1, 1, 0.1, 0.0, 1.0)
    src = np.arange(3 * 2 * 9, dtype=np.float32).reshape(2, 3, 9) + 1
    dst = np.zeros_like(src)
    obst = np.ones((2, 3), np.uint8)  # all solid: dst is a pure permutation of src
    oracle.timestep_rows(p, src, dst, obst, 0, 2, oracle.ACCUM_FLOAT, 1)
    # cell (x=0, y=0): east-going population comes from x=-1 -> x=2, north-going from y=-1 -> y=1
    pulled = [src[0, 0, 0], src[0, 2, 1], src[1, 0, 2], src[0, 1, 3], src[1, 0, 4], src[1, 2, 5], src[1, 1, 6], src[1, 1, 7],
              src[1, 2, 8]]
    assert np.array_equal(dst[0, 0], oracle.rebound_cell(np.array(pulled, np.float32)))


# ---- the reference's goldens ----------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES)
def test_oracle_av_vels_within_check_tolerance_of_shipped_goldens(case):
    """check/<case>.av_vels.dat (double-precision provenance) vs the fp32 oracle: max deviation < 1 %
    (measured 0.04-0.14 %)."""
    fx = oracle_fixture(case)
    gold = golden_av_vels(case)
    for key in ("av_f32sum", "av_f64sum"):
        d = check_tool.get_diff_values(gold, fx[key].astype(np.float64))
        assert np.isfinite(d["max_diff_pcnt"]) and abs(d["max_diff_pcnt"]) < 0.2, (key, d)


@pytest.mark.parametrize("case", ["128x128", "128x256"])
def test_oracle_final_state_within_check_tolerance_of_shipped_goldens(case):
    fx = oracle_fixture(case)
    gold = golden_final_state(case)  # columns x, y, pressure; jj outer, ii inner
    pressure = fx["pressure"].astype(np.float64).reshape(-1)
    assert int(fx["stride"]) == 1 and pressure.size == gold.shape[0]
    d = check_tool.get_diff_values(gold[:, 2], pressure)
    assert np.isfinite(d["max_diff_pcnt"]) and abs(d["max_diff_pcnt"]) < 0.1, d


@pytest.mark.parametrize("case,steps", [("128x128", 3000), ("128x256", 1500), ("256x256", 700), ("1024x1024", 40)])
def test_oracle_source_reproduces_the_committed_fixture_prefix(oracle, case, steps):
    """The fixtures were generated by the oracle and pinned byte-for-byte to the reference program; this
    re-runs today's oracle source and requires the same av_vels bits, so the two cannot drift apart."""
    p = oracle.load_params(params_path(case))
    obst = oracle.load_obstacles(obstacles_path(case), p.nx, p.ny)
    fx = oracle_fixture(case)
    assert int(fx["free_cells"]) == int((obst == 0).sum())
    _, av_f, av_d = oracle.run_both(p, oracle.init_cells(p), obst, steps)
    assert np.array_equal(av_f.view(np.uint32), fx["av_f32sum"][:steps].view(np.uint32))
    assert np.array_equal(av_d.view(np.uint32), fx["av_f64sum"][:steps].view(np.uint32))


def test_openmp_and_sequential_sweeps_agree(oracle):
    p = oracle.load_params(params_path("128x256"))
    obst = oracle.load_obstacles(obstacles_path("128x256"), p.nx, p.ny)
    a, av1 = oracle.run(p, oracle.init_cells(p), obst, 50, accum=oracle.ACCUM_DOUBLE, threads=1)
    b, av4 = oracle.run(p, oracle.init_cells(p), obst, 50, accum=oracle.ACCUM_DOUBLE, threads=4)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert ulp_distance(av1, av4).max() <= 1


# ---- the unmodified reference program (only where oracle/_ref was built) -------------------------------
@pytest.mark.slow
def test_oracle_is_byte_identical_to_the_reference_program(oracle, tmp_path):
    """oracle/_ref/compareLbm_strict = main/LastChance.cpp compiled unmodified with -ffp-contract=off.
    Same input, same text output (av_vels.dat and final_state.dat) as the restated oracle."""
    if not os.path.exists(oracle.REF_STRICT):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    pfile = tmp_path / "short.params"
    pfile.write_text("128\n128\n1500\n10\n0.1\n0.005\n1.85\n")
    ofile = obstacles_path("128x128")
    refdir, oradir = tmp_path / "ref", tmp_path / "ora"
    refdir.mkdir(), oradir.mkdir()
    subprocess.run([oracle.REF_STRICT, str(pfile), ofile], cwd=refdir, check=True, capture_output=True)
    subprocess.run([oracle.CLI_PATH, str(pfile), ofile, "--transposed-flag", "--outdir", str(oradir)], check=True,
                   capture_output=True)
    for name in ("av_vels.dat", "final_state.dat"):
        assert (refdir / name).read_bytes() == (oradir / name).read_bytes(), name


def test_fixture_records_reference_reynolds_number():
    """main/LastChance.cpp prints Reynolds 9.777...E+00 for 128x128."""
    fx = oracle_fixture("128x128")
    assert float(fx["reynolds"]) == pytest.approx(9.7773056, rel=1e-6)
    assert len(str(fx["cells_sha256"])) == 64 and hashlib.sha256(b"").hexdigest() != str(fx["cells_sha256"])


def test_reference_built_with_its_own_flags_stays_within_1e4_of_the_strict_oracle(oracle, tmp_path):
    """The reference's CMake flags allow FMA contraction (-O3 -march=native, gcc default -ffp-contract=fast); the oracle and
    the GPU kernels round every operation once.  Quantify the difference on 128x128, 2000 steps: av_vels within 1e-4
    relative (SURVEY.md section 6 measured 1.4e-5), i.e. two orders of magnitude inside check.py's 1 % rule."""
    if not os.path.exists(oracle.REF_FMA):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    pfile = tmp_path / "short.params"
    pfile.write_text("128\n128\n2000\n10\n0.1\n0.005\n1.85\n")
    subprocess.run([oracle.REF_FMA, str(pfile), obstacles_path("128x128")], cwd=tmp_path, check=True, capture_output=True)
    ref = np.loadtxt(tmp_path / "av_vels.dat", usecols=[1])
    fx = oracle_fixture("128x128")
    ours = fx["av_f32sum"][:2000].astype(np.float64)
    rel = np.abs(ref - ours) / np.abs(ours)
    assert rel.max() < 1e-4
