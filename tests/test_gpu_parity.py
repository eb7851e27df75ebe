"""GPU parity: the CUDA path, called through the C ABI (liblbm_b200.so), against the CPU oracle.

Bar (integer/bit work is n/a here; this is fp32): the lattice must be BIT-EXACT -- the kernels issue
every fp32 operation as an explicit round-to-nearest intrinsic in the reference's order and the
oracle is compiled with -ffp-contract=off.  av_vels is a sum whose order differs (GPU: fp64 tree,
oracle: fp64 row-major), so it is compared after rounding to fp32 with a tolerance of 1 ulp."""
import numpy as np
import pytest

from conftest import CASES, obstacles_path, params_path, ulp_distance

pytestmark = pytest.mark.gpu


def _random_problem(oracle, lbm, nx, ny, iters, seed, solid_frac=0.05, accel=0.005):
    rng = np.random.default_rng(seed)
    p = lbm.Params(nx, ny, iters, 10, 0.1, accel, 1.85)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, accel, 1.85)
    obst = (rng.random((ny, nx)) < solid_frac).astype(np.uint8)
    # perturbed initial state so that every population differs from cell to cell
    cells = lbm.initial_cells(p) * (1.0 + 0.1 * rng.random((ny, nx, 9))).astype(np.float32)
    return p, op, cells.astype(np.float32), obst


def _assert_matches_oracle(oracle, lbm, p, op, cells, obst, iters, **solver_kw):
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p, **solver_kw) as s:
        s.upload(cells, obst)
        s.run(iters)
        got_cells, got_av = s.download()
        assert s.free_cells == int((obst == 0).sum())
    assert got_cells.shape == want_cells.shape
    mism = np.flatnonzero(got_cells.view(np.uint32) != want_cells.view(np.uint32))
    assert mism.size == 0, f"{mism.size} lattice values differ from the oracle (first flat index {mism[0]})"
    assert np.all(np.isfinite(got_av))
    assert ulp_distance(got_av, want_av).max() <= 1
    return got_cells, got_av


@pytest.mark.parametrize("case", ["128x128", "128x256"])
def test_shipped_inputs_bit_exact_short(oracle, lbm, case):
    """vec4 kernel on the reference's own inputs, 1500 steps, bit-exact lattice."""
    p = lbm.Params.from_file(params_path(case))
    op = oracle.load_params(params_path(case))
    obst = lbm.load_obstacles(p.nx, p.ny, obstacles_path(case))
    assert np.array_equal(obst, oracle.load_obstacles(obstacles_path(case), p.nx, p.ny))
    with lbm.Solver(p) as s:
        assert s.kernel_variant == "vec4-cluster"  # 16 K / 32 K cells: the lattices live in cluster shared memory
    _assert_matches_oracle(oracle, lbm, p, op, lbm.initial_cells(p), obst, 1500)


@pytest.mark.parametrize("nx,ny", [(100, 37), (3, 2), (1, 1), (31, 5), (257, 3), (130, 64), (129, 6), (255, 7), (127, 14),
                                   (385, 15), (2, 8), (1, 23), (1000, 30), (640 + 77, 9)])
@pytest.mark.parametrize("persistent", [None, "1"])
def test_ragged_widths_hybrid_and_scalar_kernels(oracle, lbm, monkeypatch, nx, ny, persistent):
    """Widths that are not a multiple of the tile (nx % 128 != 0): float4 path on the full tiles + scalar path on the
    last tile column ("hybrid"), scalar only below 128; including the reference KAT shape 3x2 and 1x1; last
    tile columns with 1, 2 and 127 cells; slab heights that put the north ghost row in a tile of its own.  Through the
    per-step kernels (default; 40 steps = two graph replays + 8 plain launches) and the persistent cooperative kernel."""
    if persistent:
        monkeypatch.setenv("LBM_B200_PERSISTENT", persistent)
    p, op, cells, obst = _random_problem(oracle, lbm, nx, ny, 40, seed=nx * 1000 + ny)
    with lbm.Solver(p) as s:
        assert s.kernel_variant == ("hybrid" if nx > 128 else "scalar") + ("-persistent" if persistent else "")
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 40)


@pytest.mark.parametrize("nx,ny", [(128, 1), (128, 2), (256, 7), (1024, 33), (2048, 16), (384, 9), (128, 6), (256, 14),
                                   (128, 15), (512, 2500)])
def test_vec4_kernel_random(oracle, lbm, nx, ny):
    """float4 path (nx % 128 == 0) with random solids everywhere (also on the driven row) and a perturbed
    lattice; one and several tile columns, partial last tile rows, ghost row alone in the last tile row."""
    p, op, cells, obst = _random_problem(oracle, lbm, nx, ny, 30, seed=nx + ny, solid_frac=0.1)
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 30)


@pytest.mark.parametrize("nx,ny", [(128, 3), (256, 64), (1024, 9), (384, 20)])
def test_scalar_kernel_on_vectorisable_sizes(oracle, lbm, nx, ny, monkeypatch):
    """The any-width kernel forced onto sizes the float4 kernel normally takes: both must give the same bits."""
    monkeypatch.setenv("LBM_B200_KERNEL", "scalar")
    p, op, cells, obst = _random_problem(oracle, lbm, nx, ny, 30, seed=nx * 7 + ny, solid_frac=0.1)
    with lbm.Solver(p) as s:
        assert s.kernel_variant == "scalar"
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 30)
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 30, devices=[0, 0])


@pytest.mark.parametrize("nx,ny", [(256, 64), (1024, 33), (300, 21), (77, 19), (512, 2500)])
def test_per_step_kernels_on_small_grids(oracle, lbm, nx, ny, monkeypatch):
    """Small lattices normally take the persistent cooperative kernel; LBM_B200_PERSISTENT=0 sends them through the
    one-launch-per-step kernels (programmatic dependent launch, single-level reduction) that big grids use."""
    monkeypatch.setenv("LBM_B200_PERSISTENT", "0")
    p, op, cells, obst = _random_problem(oracle, lbm, nx, ny, 25, seed=nx + 11 * ny, solid_frac=0.07)
    with lbm.Solver(p) as s:
        assert "persistent" not in s.kernel_variant
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 25)


def test_persistent_kernel_with_several_tiles_per_cta(oracle, lbm, monkeypatch):
    """LBM_B200_PERSISTENT=1 forces the persistent kernel onto a lattice with more tiles than resident CTAs."""
    monkeypatch.setenv("LBM_B200_PERSISTENT", "1")
    p, op, cells, obst = _random_problem(oracle, lbm, 2048, 700, 9, seed=5, solid_frac=0.03)
    with lbm.Solver(p) as s:
        assert s.kernel_variant == "vec4-persistent"
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 9)


def test_run_in_pieces_equals_one_run(oracle, lbm):
    p, op, cells, obst = _random_problem(oracle, lbm, 256, 64, 50, seed=7)
    want_cells, want_av = oracle.run(op, cells, obst, 50, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p) as s:
        s.upload(cells, obst)
        for n in (1, 2, 17, 30):
            s.run(n)
        got_cells, got_av = s.download()
        assert s.steps_done == 50
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1
    # odd number of steps: the state lives in the other buffer
    want_cells, _ = oracle.run(op, cells, obst, 33, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p) as s:
        s.upload(cells, obst)
        s.run(33)
        got_cells, _ = s.download()
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))


@pytest.mark.parametrize("slabs", [2, 3, 4])
@pytest.mark.parametrize("nx,ny", [(256, 64), (100, 19), (128, 13), (384, 31)])
def test_multi_slab_on_one_gpu_bit_exact(oracle, lbm, slabs, nx, ny):
    """Row-slab decomposition with writer-filled ghost rows and epoch flags, all slabs on GPU 0 (kernels
    run back to back on one stream).  The lattice must not depend on the decomposition."""
    p, op, cells, obst = _random_problem(oracle, lbm, nx, ny, 25, seed=slabs)
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 25, devices=[0] * slabs)


def test_synthetic_channel_matches_host_twin(oracle, lbm):
    """lbm_init_synthetic == uploading the numpy twin of the same mask + Cells::initialise."""
    nx, ny, iters = 1024, 96, 12
    p = lbm.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    obst = lbm.synthetic_obstacles(nx, ny, seed=1234, solid_prob=0.01)
    assert obst[0].all() and obst[-1].all() and not obst[ny - 2].any()
    assert 0.005 < obst[1:-2].mean() < 0.015
    want_cells, want_av = oracle.run(op, lbm.initial_cells(p), obst, iters, accum=oracle.ACCUM_DOUBLE)
    with lbm.Solver(p) as s:
        s.init_synthetic(1234, 0.01)
        assert s.free_cells == int((obst == 0).sum())
        s.run(iters)
        got_cells, got_av = s.download()
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1


@pytest.mark.parametrize("nx,ny,devices", [(256, 40, None), (100, 19, None), (384, 31, [0, 0, 0])])
def test_device_final_state_fields_equal_host_fields(oracle, lbm, nx, ny, devices):
    """lbm_download_fields (u_x, u_y, u, pressure evaluated on the GPU) == the oracle's / host writer's fields, bit
    for bit, obstacle cells included (u = 0, pressure = density/3)."""
    p, op, cells, obst = _random_problem(oracle, lbm, nx, ny, 20, seed=nx + 3 * ny, solid_frac=0.1)
    kw = {"devices": devices} if devices else {}
    with lbm.Solver(p, **kw) as s:
        s.upload(cells, obst)
        s.run(20)
        got_cells, _ = s.download(av_vels=False)
        fields = s.download_fields()
    ux, uy, u, pr = oracle.final_state_fields(op, got_cells, obst)
    want = np.stack([ux, uy, u, pr], axis=-1).astype(np.float32)
    assert fields.shape == want.shape
    assert np.array_equal(fields.view(np.uint32), want.view(np.uint32))


def test_division_and_sqrt_helpers_match_ieee_library(lbm):
    """The kernels replace __fdiv_rn/__fsqrt_rn by the library's own fast-path sequence under a wider
    guard (zero numerators, sqrt(0)).  Exhaustive for rho/9, rho/36 and sqrt; 2^30 random pairs for a/b."""
    r = lbm.selftest_math(0, 1 << 30)
    assert r["div9_n"] == (0x5D800000 - 0x21800000 + 1)
    assert r["sqrt_n"] > 1.9e9 and r["div_n"] > 0.9 * (1 << 30)
    assert r["div9_bad"] == 0 and r["div36_bad"] == 0 and r["sqrt_bad"] == 0 and r["div_bad"] == 0


@pytest.mark.parametrize("density,accel,omega", [(0.1, 0.0, 1.85), (1.0, 0.01, 0.6), (0.05, 0.002, 1.99), (0.3, 0.02, 1.0)])
@pytest.mark.parametrize("nx,ny,env", [(256, 24, None), (200, 17, "1"), (256, 24, "1"), (256, 64, None)])
def test_other_physical_parameters(oracle, lbm, monkeypatch, density, accel, omega, nx, ny, env):
    """Relaxation from under- to almost marginally stable over-relaxation, no forcing, strong forcing (the guard of the
    first accelerate then refuses some cells), other densities: still bit-exact -- per-step kernels (default), the
    persistent cooperative kernel (LBM_B200_PERSISTENT=1) and the cluster kernel (256 x 64)."""
    if env is not None:
        monkeypatch.setenv("LBM_B200_PERSISTENT", env)
    rng = np.random.default_rng(int(omega * 100) + nx)
    p = lbm.Params(nx, ny, 30, 10, density, accel, omega)
    op = oracle.make_params(nx, ny, 30, 10, density, accel, omega)
    obst = (rng.random((ny, nx)) < 0.06).astype(np.uint8)
    cells = (lbm.initial_cells(p) * (1.0 + 0.2 * rng.random((ny, nx, 9))).astype(np.float32)).astype(np.float32)
    cells[ny - 2, ::7, 3] = 0.0  # f3 - w1 <= 0: the guarded first accelerate must skip these cells
    _assert_matches_oracle(oracle, lbm, p, op, cells, obst, 30)


def test_zero_steps_and_download_of_the_initial_state(oracle, lbm):
    p, op, cells, obst = _random_problem(oracle, lbm, 256, 16, 4, seed=2)
    with lbm.Solver(p) as s:
        s.upload(cells, obst)
        s.run(0)  # applies the one-off accelerate_flow only
        got, av = s.download()
        assert s.steps_done == 0 and av.size == 0
    want = cells.copy()
    oracle.accelerate_flow(op, want, obst)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_slab_diagnostics(lbm):
    p = lbm.Params(256, 50, 1, 10, 0.1, 0.005, 1.85)
    obst = np.zeros((50, 256), np.uint8)
    obst[:10] = 1
    with lbm.Solver(p, devices=[0, 0, 0]) as s:
        assert [d["free_cells"] for d in s.slabs] == [-1, -1, -1]
        s.upload(lbm.initial_cells(p), obst)
        info = s.slabs
    assert [(d["row_begin"], d["row_count"]) for d in info] == [(0, 16), (16, 17), (33, 17)]
    assert [d["free_cells"] for d in info] == [6 * 256, 17 * 256, 17 * 256] and all(d["device"] == 0 for d in info)


def test_errors_are_reported_not_thrown(lbm):
    p = lbm.Params(128, 128, 10, 10, 0.1, 0.005, 1.85)
    with lbm.Solver(p) as s:
        with pytest.raises(lbm.LbmError) as e:
            s.run(1)  # before upload
        assert e.value.code == -3
        s.upload(lbm.initial_cells(p), np.zeros((128, 128), np.uint8))
        with pytest.raises(lbm.LbmError) as e:
            s.run(11)  # beyond max_iters
        assert e.value.code == -1
    with pytest.raises(lbm.LbmError):
        lbm.Solver(lbm.Params(0, 5, 1, 1, 0.1, 0.005, 1.85))
    with pytest.raises(lbm.LbmError):
        lbm.Solver(p, num_gpus=1000)
