"""lbm_solve_host: the whole job (upload, K timesteps, download) as one call that solves the grid band by band so that
PCIe transfers in both directions and the sweeps overlap.  The result must be the sequential job's: lattice bit-identical
to the oracle, av_vels within 1 fp32 ulp -- whatever the band height, wherever the driven row falls."""
import numpy as np
import pytest

from conftest import ulp_distance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,ny,iters,band", [(256, 256, 5, 64),     # four bands, margin 7
                                              (384, 520, 12, 128),   # margin 15; last band 8 rows
                                              (128, 512, 7, 64),     # eight bands over three workers
                                              (200, 256, 3, 64),     # ragged width: hybrid kernel in the band solvers
                                              (256, 256, 0, 64),     # zero steps: the one-off accelerate only
                                              (256, 1024, 30, 256)])  # margin 31, graph-replayed per-step loop inside a band
def test_banded_solve_equals_the_oracle(oracle, lbm, monkeypatch, nx, ny, iters, band):
    monkeypatch.setenv("LBM_B200_BAND_ROWS", str(band))
    p, cells, obst = lbm.perturbed_problem(nx, ny, max(iters, 1), seed=nx + 3 * ny + iters)
    op = oracle.make_params(nx, ny, max(iters, 1), 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    got_cells, got_av, ms = lbm.solve_host(p, cells, obst, iters)
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert got_av.shape == want_av.shape and (iters == 0 or ulp_distance(got_av, want_av).max() <= 1)
    assert ms > 0


def test_small_or_long_jobs_take_the_sequential_path(oracle, lbm):
    """No band override: a 256 x 256 grid is far below the band size, so the three calls run internally; same answer."""
    p, cells, obst = lbm.perturbed_problem(256, 256, 40, seed=17)
    op = oracle.make_params(256, 256, 40, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, 40, accum=oracle.ACCUM_DOUBLE)
    got_cells, got_av, _ = lbm.solve_host(p, cells, obst)
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1


def test_banded_solve_of_the_shipped_input_matches_the_plain_calls(lbm, monkeypatch):
    """The reference's 1024 x 1024 input (box walls, a solid column), 60 steps in 128-row bands vs upload/run/download."""
    from conftest import obstacles_path, params_path
    import dataclasses
    p = dataclasses.replace(lbm.Params.from_file(params_path("1024x1024")), max_iters=60)
    obst = lbm.load_obstacles(p.nx, p.ny, obstacles_path("1024x1024"))
    cells = lbm.initial_cells(p)
    with lbm.Solver(p) as s:
        s.upload(cells, obst)
        s.run(60)
        want_cells, want_av = s.download()
    monkeypatch.setenv("LBM_B200_BAND_ROWS", "256")  # 60 steps need a margin of 63 rows: too long a job for 256-row bands -> sequential
    got_cells, got_av, _ = lbm.solve_host(p, cells, obst, 60)
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    p10 = dataclasses.replace(p, max_iters=10)
    with lbm.Solver(p10) as s:
        s.upload(cells, obst)
        s.run(10)
        want_cells, want_av = s.download()
    got_cells, got_av, _ = lbm.solve_host(p10, cells, obst, 10)  # margin 15, bands of 256 rows: banded
    assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
    assert ulp_distance(got_av, want_av).max() <= 1


def test_pipeline_is_reusable(oracle, lbm, monkeypatch):
    """lbm_pipeline_create once, lbm_pipeline_solve on two different inputs of the same shape."""
    monkeypatch.setenv("LBM_B200_BAND_ROWS", "64")
    nx, ny, iters = 256, 320, 6
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    with lbm.Pipeline(lbm.Params(nx, ny, iters, 10, 0.1, 0.005, 1.85)) as pipe:
        assert pipe.info["bands"] == 5 and pipe.info["band_rows"] == 64 and pipe.info["margin_rows"] == 7
        for seed in (1, 2):
            _, cells, obst = lbm.perturbed_problem(nx, ny, iters, seed)
            want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
            got_cells, got_av = pipe.solve(cells, obst)
            assert np.array_equal(got_cells.view(np.uint32), want_cells.view(np.uint32))
            assert ulp_distance(got_av, want_av).max() <= 1
        assert pipe.info["last_ms"] > 0
        with pytest.raises(lbm.LbmError, match="alias"):
            pipe.solve_raw(cells.ctypes.data, obst.ctypes.data, cells.ctypes.data, None)


@pytest.mark.parametrize("nx,ny,parts,iters,band", [(256, 512, 4, 5, None),    # 128-row windows, default band sizing: 56 + 56 + 16 rows
                                                    (256, 512, 4, 5, "32"),
                                                    (384, 256, 2, 9, None),    # margin 15: the two windows see each other's rows twice over
                                                    (128, 192, 3, 3, "64"),    # one band per window
                                                    (200, 128, 1, 4, "32"),    # a single window that is the whole grid (wraps onto itself)
                                                    (64, 64, 1, 20, None),     # margin 23 on a 64-row grid: bands shrink to 16 rows so
                                                    (128, 96, 2, 25, None)])   # that no band holds the driven row twice (test_pipeline_model.py)
def test_window_pipelines_reproduce_the_whole_grid(oracle, lbm, monkeypatch, nx, ny, parts, iters, band):
    """lbm_pipeline_create_window: every 'rank' solves its slab from the slab's initial rows plus margin rows of its neighbours,
    independently of the others; together they must give the single-domain oracle's lattice and av_vels."""
    if band:
        monkeypatch.setenv("LBM_B200_BAND_ROWS", band)
    p, cells, obst = lbm.perturbed_problem(nx, ny, iters, seed=nx + ny + parts)
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want_cells, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE)
    tot, free = np.zeros(iters), 0
    for rank in range(parts):  # rank order: the documented order of the last reduction stage
        r0, n = rank * ny // parts, (rank + 1) * ny // parts - rank * ny // parts
        with lbm.Pipeline(p, iters, window=(r0, n)) as pipe:
            got, t, f = pipe.solve_window(cells, obst)
        assert np.array_equal(got.view(np.uint32), want_cells[r0:r0 + n].view(np.uint32)), rank
        tot, free = tot + t, free + f
    assert free == int((obst == 0).sum())
    av = (tot.astype(np.float32) / np.float32(free)).astype(np.float32)
    assert ulp_distance(av, want_av).max() <= 1


def test_window_argument_errors(lbm):
    p = lbm.Params(256, 256, 4, 10, 0.1, 0.005, 1.85)
    with pytest.raises(lbm.LbmError, match="multiple of 8"):
        lbm.Pipeline(p, 4, window=(0, 100))
    with pytest.raises(lbm.LbmError, match="not inside"):
        lbm.Pipeline(p, 4, window=(200, 128))
    with pytest.raises(lbm.LbmError, match="too short"):
        lbm.Pipeline(lbm.Params(64, 40, 20, 10, 0.1, 0.005, 1.85), 20, window=(0, 40))
    with lbm.Pipeline(p, 4, window=(0, 128)) as pipe, pytest.raises(lbm.LbmError, match="window pipeline"):
        pipe.solve_raw(1, 1, 2, None)
