"""Host-side problem I/O (liblbm_host.so = lbm-graphcore_b200/csrc/lbm_host.hpp, the mirror of the reference's
LbmParams.hpp / LatticeBoltzmannUtils.hpp) against the oracle's loaders and writers and the numpy twin in
problem.py.  CPU only."""
import ctypes
import os

import numpy as np
import pytest

from conftest import CASES, obstacles_path, params_path


@pytest.fixture(scope="module")
def host(lbm):
    L = ctypes.CDLL(lbm.HOST_LIB_PATH)
    ll, f, vp = ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
    P = ctypes.POINTER
    L.lbm_host_load_params.argtypes = [ctypes.c_char_p, P(ll), P(ll), P(ll), P(ll), P(f), P(f), P(f)]
    L.lbm_host_load_obstacles.argtypes = [ctypes.c_char_p, ll, ll, vp]
    L.lbm_host_load_obstacles.restype = ll
    L.lbm_host_init_cells.argtypes = [ll, ll, f, vp]
    L.lbm_host_average_velocity.argtypes = [ll, ll, vp, vp]
    L.lbm_host_average_velocity.restype = f
    L.lbm_host_total_density.argtypes = [ll, ll, vp]
    L.lbm_host_total_density.restype = f
    L.lbm_host_reynolds_number.argtypes = [f, ll, f]
    L.lbm_host_reynolds_number.restype = f
    L.lbm_host_write_av_vels.argtypes = [ctypes.c_char_p, vp, ll]
    L.lbm_host_write_final_state.argtypes = [ctypes.c_char_p, ll, ll, f, vp, vp]
    L.lbm_host_write_final_state_fields.argtypes = [ctypes.c_char_p, ll, ll, vp, vp, ctypes.c_int]
    L.lbm_host_synthetic_obstacles.argtypes = [ll, ll, ctypes.c_uint64, f, ll, ll, vp]
    return L


def load_params(host, path):
    ll, f = ctypes.c_longlong, ctypes.c_float
    v = [ll(), ll(), ll(), ll(), f(), f(), f()]
    rc = host.lbm_host_load_params(path.encode(), *[ctypes.byref(x) for x in v])
    return rc, [x.value for x in v]


@pytest.mark.parametrize("case", CASES)
def test_loaders_agree_with_oracle_and_numpy_twin(host, lbm, oracle, case):
    rc, vals = load_params(host, params_path(case))
    assert rc == 0
    op = oracle.load_params(params_path(case))
    p = lbm.Params.from_file(params_path(case))
    assert vals[:4] == [op.nx, op.ny, op.max_iters, op.reynolds_dim] == [p.nx, p.ny, p.max_iters, p.reynolds_dim]
    assert np.float32(vals[4]) == np.float32(op.density) == np.float32(p.density)
    assert np.float32(vals[5]) == np.float32(op.accel) and np.float32(vals[6]) == np.float32(op.omega)
    obst = np.empty((p.ny, p.nx), np.uint8)
    solids = host.lbm_host_load_obstacles(obstacles_path(case).encode(), p.nx, p.ny, obst.ctypes.data)
    want = oracle.load_obstacles(obstacles_path(case), p.nx, p.ny)
    assert np.array_equal(obst, want) and np.array_equal(obst, lbm.load_obstacles(p.nx, p.ny, obstacles_path(case)))
    assert solids == int(want.sum())  # duplicate lines in the shipped files are idempotent
    cells = np.empty((p.ny, p.nx, 9), np.float32)
    assert host.lbm_host_init_cells(p.nx, p.ny, p.density, cells.ctypes.data) == 0
    assert np.array_equal(cells.view(np.uint32), oracle.init_cells(op).view(np.uint32))
    assert np.array_equal(cells.view(np.uint32), lbm.initial_cells(p).view(np.uint32))


def test_bad_inputs_are_rejected(host, tmp_path):
    assert load_params(host, str(tmp_path / "missing.params"))[0] != 0
    short = tmp_path / "short.params"
    short.write_text("128\n128\n100\n")
    assert load_params(host, str(short))[0] != 0
    neg = tmp_path / "neg.params"
    neg.write_text("-1\n128\n100\n10\n0.1\n0.005\n1.85\n")
    assert load_params(host, str(neg))[0] != 0
    obst = np.zeros((4, 4), np.uint8)
    oob = tmp_path / "oob.dat"
    oob.write_text("1 1 1\n9 0 1\n")  # the reference writes out of bounds here (LbmParams.hpp:115); we refuse
    assert host.lbm_host_load_obstacles(str(oob).encode(), 4, 4, obst.ctypes.data) < 0
    # like the reference (LbmParams.hpp:101-106) a line whose third field is not 1 ENDS parsing with a message;
    # what was read before it is kept
    bad = tmp_path / "bad.dat"
    bad.write_text("2 2 1\n1 1 2\n3 3 1\n")
    assert host.lbm_host_load_obstacles(str(bad).encode(), 4, 4, obst.ctypes.data) == 1 and obst[2, 2] == 1 and obst[3, 3] == 0
    ok = tmp_path / "ok.dat"
    ok.write_text("1 1 1\n1 1 1\n3 0 1\n")
    assert host.lbm_host_load_obstacles(str(ok).encode(), 4, 4, obst.ctypes.data) == 2 and obst[1, 1] == 1 and obst[0, 3] == 1


def test_writers_match_the_oracle_numerically_and_the_reference_format(host, oracle, tmp_path):
    """av_vels.dat: '<i>:\\t<%.12e>'; final_state.dat: 'ii jj u_x u_y u pressure obstacle' (jj outer), the
    format of LatticeBoltzmannUtils.hpp:208-281.  Values must equal the oracle's (same fp32 arithmetic)."""
    case = "128x128"
    p = oracle.load_params(params_path(case))
    obst = oracle.load_obstacles(obstacles_path(case), p.nx, p.ny)
    cells, av = oracle.run(p, oracle.init_cells(p), obst, 60, accum=oracle.ACCUM_FLOAT)
    a, b = str(tmp_path / "av_vels.dat"), str(tmp_path / "final_state.dat")
    assert host.lbm_host_write_av_vels(a.encode(), av.ctypes.data, av.size) == 0
    assert host.lbm_host_write_final_state(b.encode(), p.nx, p.ny, p.density, cells.ctypes.data, obst.ctypes.data) == 0
    lines = open(a).read().splitlines()
    assert len(lines) == 60 and lines[7].startswith("7:\t") and "e-" in lines[7].lower()
    assert np.array_equal(np.loadtxt(a, usecols=[1]).astype(np.float32), av)
    fs = np.loadtxt(b)
    assert fs.shape == (p.nx * p.ny, 7)
    assert np.array_equal(fs[:, 0], np.tile(np.arange(p.nx), p.ny)) and np.array_equal(fs[:, 1], np.repeat(np.arange(p.ny), p.nx))
    ux, uy, u, pr = oracle.final_state_fields(p, cells, obst)
    for col, want in ((2, ux), (3, uy), (4, u), (5, pr)):
        got = fs[:, col].astype(np.float32).reshape(p.ny, p.nx)
        assert np.array_equal(got, want), col
    assert np.array_equal(fs[:, 6].reshape(p.ny, p.nx), obst)
    # the oracle's own writer (reference C format, "%.12E") parses to the same numbers
    c = str(tmp_path / "fs_oracle.dat")
    oracle.write_final_state(c, p, cells, obst)
    assert np.array_equal(np.loadtxt(c)[:, :6], fs[:, :6])


def test_scalar_helpers(host, lbm, oracle):
    p = oracle.make_params(64, 32, 1, 10, 0.1, 0.005, 1.85)
    rng = np.random.default_rng(3)
    cells = (oracle.init_cells(p) * (1 + 0.2 * rng.random((32, 64, 9)))).astype(np.float32)
    obst = (rng.random((32, 64)) < 0.1).astype(np.uint8)
    av = host.lbm_host_average_velocity(64, 32, cells.ctypes.data, obst.ctypes.data)
    assert av == pytest.approx(oracle.av_velocity(p, cells, obst), rel=2e-6)
    assert host.lbm_host_total_density(64, 32, cells.ctypes.data) == pytest.approx(oracle.total_density(p, cells), rel=1e-5)
    assert host.lbm_host_reynolds_number(1.85, 10, 0.01) == pytest.approx(oracle.reynolds(p, 0.01), rel=1e-6)
    assert lbm.Params(64, 32, 1, 10, 0.1, 0.005, 1.85).reynolds_number(0.01) == pytest.approx(oracle.reynolds(p, 0.01), rel=1e-6)


def test_synthetic_channel_host_and_numpy_twins_agree(host, lbm):
    nx, ny = 200, 77
    full = lbm.synthetic_obstacles(nx, ny, 1234, 0.01)
    got = np.empty((20, nx), np.uint8)
    assert host.lbm_host_synthetic_obstacles(nx, ny, 1234, 0.01, 50, 20, got.ctypes.data) == 0
    assert np.array_equal(got, full[50:70])
    assert np.array_equal(lbm.synthetic_obstacles(nx, ny, 1234, 0.01, 50, 20), full[50:70])
    assert full[0].all() and full[-1].all() and not full[ny - 2].any()


def test_fields_writer_and_threaded_formatting_give_the_same_bytes(host, oracle, tmp_path):
    """final_state.dat from device-style fields (u_x, u_y, u, pressure per cell) on 1, 3 and all threads must be
    byte-identical to the cells-based writer (the writer, not the solve, dominates 1024^2)."""
    p = oracle.make_params(300, 70, 1, 10, 0.1, 0.005, 1.85)
    rng = np.random.default_rng(5)
    cells = (oracle.init_cells(p) * (1 + 0.3 * rng.random((70, 300, 9)))).astype(np.float32)
    obst = (rng.random((70, 300)) < 0.1).astype(np.uint8)
    ref = str(tmp_path / "cells.dat")
    assert host.lbm_host_write_final_state(ref.encode(), 300, 70, 0.1, cells.ctypes.data, obst.ctypes.data) == 0
    ux, uy, u, pr = oracle.final_state_fields(p, cells, obst)
    fields = np.ascontiguousarray(np.stack([ux, uy, u, pr], axis=-1), dtype=np.float32)
    want = open(ref, "rb").read()
    assert want.count(b"\n") == 300 * 70
    for threads in (1, 3, 0):
        out = str(tmp_path / f"fields{threads}.dat")
        assert host.lbm_host_write_final_state_fields(out.encode(), 300, 70, fields.ctypes.data, obst.ctypes.data, threads) == 0
        assert open(out, "rb").read() == want
